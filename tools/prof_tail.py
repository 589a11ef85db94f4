"""Time one forward (CUDA events, rotating inputs) -- used to sweep tail-kernel knobs through environment variables."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vad_b200 import synthetic as S
from vad_b200.engine import VadEngine
B, T = int(os.environ.get("B", 256)), int(os.environ.get("T", 512))
eng = VadEngine.from_state_dict(S.random_state(0, 64, 3, 128), compute_dtype="bf16")
g = torch.Generator().manual_seed(0)
xs = [(torch.randn(B, T, 64, generator=g) * 2 - 3).cuda().to(torch.bfloat16) for _ in range(6)]
for i in range(5):
    eng.forward(xs[i % 6], want_logp=False)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(50):
        eng.forward(xs[i % 6], want_logp=False)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 50)
print(f"{os.environ.get('TAG','')} forward {best*1e3:.1f} us")
