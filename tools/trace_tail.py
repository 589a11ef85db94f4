"""One warm forward with VADB_TAIL_TRACE set (1: q|k|v variant, 0: last-layer variant): the layer-tail kernel prints the
clock64 stamps of CTA 0's third tile to stderr (developer tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vad_b200 import synthetic as S
from vad_b200.engine import VadEngine
eng = VadEngine.from_state_dict(S.random_state(0, 64, 3, 128), compute_dtype="bf16")
x = (torch.randn(256, 512, 64) * 2 - 3).cuda().to(torch.bfloat16)
for _ in range(2):
    eng.forward(x, want_logp=False)
torch.cuda.synchronize()
