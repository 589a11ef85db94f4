"""Run individual attention shapes in subprocesses (a trap poisons the CUDA context)."""
import subprocess, sys, os
cases = [(1, 64), (1, 128), (1, 256), (2, 512), (3, 300)]
code = r'''
import sys, torch, numpy as np
sys.path.insert(0, %r)
from vad_b200 import synthetic as S
from vad_b200.engine import VadEngine
B, T = %d, %d
eng = VadEngine.from_state_dict(S.random_state(0, 64, 3, 128), compute_dtype="bf16")
g = torch.Generator().manual_seed(0)
q, k, v = (torch.randn(B, T, 128, generator=g).cuda().to(torch.bfloat16) for _ in range(3))
o = eng.attention(q, k, v)
torch.cuda.synchronize()
s = (q.double() @ k.double().transpose(1, 2)) / np.sqrt(128.0)
want = torch.softmax(s, -1) @ v.double()
print("B=%%d T=%%d max err %%.3e" %% (B, T, (o.double() - want).abs().max().item()))
'''
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for B, T in cases:
    r = subprocess.run([sys.executable, "-c", code % (root, B, T)], capture_output=True, text=True, timeout=120)
    out = (r.stdout + r.stderr).strip().splitlines()
    keep = [l for l in out if "max err" in l or "timeout" in l or "Error" in l]
    print(f"case B={B} T={T}: rc={r.returncode} ::", " | ".join(keep[:3]))
