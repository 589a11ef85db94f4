"""Checkpoint / config loading for the inference path.

Replaces the ``torch.load`` + ``OmegaConf.create(checkpoint["config"])`` pair of
VADFromScratchPredictor.from_checkpoint (vad/predictor.py:264-267) without omegaconf: the
checkpoint written by ModelCheckpointer (vad/training/checkpointers/model_checkpointer.py:97-110)
embeds the config as a plain container, read here into an attribute-style view.
"""
from __future__ import annotations

import _codecs
from pathlib import Path
from typing import Any

import numpy as np
import torch


class Config(dict):
    """dict with attribute access (``cfg.model.self_attention.num_layers``), enough of the
    OmegaConf surface for the keys the inference path reads."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(name) from e

    def __setattr__(self, name, value):
        self[name] = value

    @classmethod
    def wrap(cls, obj: Any):
        if isinstance(obj, dict):
            return cls({k: cls.wrap(v) for k, v in obj.items()})
        if isinstance(obj, (list, tuple)):
            return type(obj)(cls.wrap(v) for v in obj)
        return obj


def _safe_globals():
    # the reference checkpoint pickles numpy scalars inside "metrics"; torch >= 2.6 defaults to
    # weights_only=True and needs these allow-listed
    # numpy >= 1.26 exposes the implementation module as np._core, older releases only as np.core
    core = getattr(np, "_core", None) or np.core
    scalar = core.multiarray.scalar
    return [(scalar, "numpy.core.multiarray.scalar"),     # pickles written by numpy < 2
            scalar, np.dtype,
            _codecs.encode, type(np.dtype("float64")), type(np.dtype("float32")),
            type(np.dtype("int64"))]


def load_checkpoint(path, map_location="cpu") -> dict:
    with torch.serialization.safe_globals(_safe_globals()):
        return torch.load(str(Path(path)), map_location=map_location)


def context_window_frames(half: int, jump: int) -> int:
    """vad/predictor.py:57-59, :270-275."""
    return 2 * (half - 1) // jump + 3
