"""Run a few forwards at the bench shape (for the ncu launch list)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vad_b200 import synthetic as S
from vad_b200.engine import VadEngine

B, T = int(os.environ.get("B", 256)), int(os.environ.get("T", 512))
dtype = os.environ.get("DTYPE", "bf16")
eng = VadEngine.from_state_dict(S.random_state(0, 64, 3, 128), compute_dtype=dtype)
g = torch.Generator().manual_seed(0)
xs = [(torch.randn(B, T, 64, generator=g) * 2 - 3).cuda().to(torch.bfloat16 if dtype == "bf16" else torch.float32) for _ in range(3)]
for i in range(int(os.environ.get("ITERS", 4))):
    eng.forward(xs[i % 3], want_logp=False)
torch.cuda.synchronize()
print("done")
