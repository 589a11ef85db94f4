from vad_b200.cli import predict_vad_from_scratch  # noqa: F401
