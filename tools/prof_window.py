"""Time the reference-Predictor-shaped window path (vad/predictor.py:159-262): L frames of F=80 log-mel,
half=19, jump=9 -> W=7-frame windows, device-resident and host-to-host."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from vad_b200 import synthetic as S
from vad_b200.engine import VadEngine

L = int(os.environ.get("L", 60000))          # 10 minutes of audio at 100 frames/s
iters = int(os.environ.get("ITERS", 10))
st = S.random_state(3, 80, 3, 128)
eng = VadEngine.from_state_dict(st, compute_dtype=os.environ.get("DTYPE", "bf16"))
feat = (torch.randn(L, 80, generator=torch.Generator().manual_seed(0)) * 2 - 3)
fd = feat.cuda()
fh = feat.numpy()
for _ in range(3):
    eng.predict_probabilities(fd, 19, 9)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    eng.predict_probabilities(fd, 19, 9)
e1.record(); torch.cuda.synchronize()
dev_ms = e0.elapsed_time(e1) / iters
for _ in range(2):
    eng.predict_probabilities(fh, 19, 9)
t0 = time.perf_counter()
for _ in range(iters):
    eng.predict_probabilities(fh, 19, 9)
host_ms = (time.perf_counter() - t0) / iters * 1e3
n = L - 38
print(f"window path L={L} ({L/100:.0f} s of audio, {n} windows, {n*7} window frames) small_attn={os.environ.get('VADB_ATTN_SMALL','1')}: "
      f"device {dev_ms:.3f} ms ({n*7/dev_ms/1e3:.1f} M window-frames/s), host-to-host {host_ms:.3f} ms "
      f"({L/100/(host_ms/1e3):.0f}x real time)")
