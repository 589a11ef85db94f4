"""Importable alias of the product package.

The package directory is named ``voice-activity-detection_b200`` (the repository layout the
build contract asks for), which is not a valid Python identifier; this shim makes it importable
as ``vad_b200`` by pointing the package search path at that directory, so
``import vad_b200.engine`` loads ``voice-activity-detection_b200/engine.py``.
"""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                          "voice-activity-detection_b200")]
__version__ = "0.1.0"
