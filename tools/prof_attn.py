"""Run the attention kernel alone at the bench shape (for ncu captures / quick timing)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vad_b200 import synthetic as S
from vad_b200.engine import VadEngine

B, T = int(os.environ.get("B", 256)), int(os.environ.get("T", 512))
iters = int(os.environ.get("ITERS", 5))
eng = VadEngine.from_state_dict(S.random_state(0, 64, 3, 128), compute_dtype="bf16")
g = torch.Generator().manual_seed(0)
sets = [tuple(torch.randn(B, T, 128, generator=g).cuda().to(torch.bfloat16) for _ in range(3)) for _ in range(3)]
for i in range(3):
    eng.attention(*sets[i % 3])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(iters):
    eng.attention(*sets[i % 3])
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / iters * 1e3
print(f"attention B={B} T={T}: {us:.1f} us/launch, {4*B*T*128*2/us/1e3:.0f} GB/s algorithmic, {4*B*T*T*128/us/1e6:.0f} TFLOP/s")
