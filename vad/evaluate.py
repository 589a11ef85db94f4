from vad_b200.cli import evaluate_vad_from_scratch  # noqa: F401
