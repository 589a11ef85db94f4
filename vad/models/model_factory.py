from vad_b200.model import ModelName, create_model  # noqa: F401
