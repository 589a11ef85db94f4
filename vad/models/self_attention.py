from vad_b200.model import SelfAttentiveVAD  # noqa: F401
