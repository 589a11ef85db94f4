"""vad_b200 -- B200-native (sm_100a) Self-Attentive VAD inference hot path.

Host-side mirror of the reference interface for this one path (``vad.predictor``,
``vad.models.self_attention``) over the C ABI of ``libvadb200.so`` (include/vadb200.h).
There is no CPU fallback: every compute entry point needs the CUDA library and a B200.
"""
__version__ = "0.1.0"
