"""Time and check every compiled softmax-schedule variant of the attention kernel (developer tool).
Each variant runs in its own subprocess (VADB_ATTN_VARIANT is read once per process)."""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os, torch, numpy as np
sys.path.insert(0, %r)
from vad_b200 import synthetic as S
from vad_b200.engine import VadEngine
eng = VadEngine.from_state_dict(S.random_state(0, 64, 3, 128), compute_dtype="bf16")
g = torch.Generator().manual_seed(0)
worst = 0.0
for (B, T, lens) in [(1, 64, None), (2, 512, None), (3, 300, [300, 17, 129]), (2, 1024, [1024, 700])]:
    q, k, v = (torch.randn(B, T, 128, generator=g).cuda().to(torch.bfloat16) for _ in range(3))
    q = q * 2.0
    ln = torch.tensor(lens, dtype=torch.int32).cuda() if lens else None
    o = eng.attention(q, k, v, ln) if ln is not None else eng.attention(q, k, v)
    torch.cuda.synchronize()
    s = (q.double() @ k.double().transpose(1, 2)) / np.sqrt(128.0)
    if lens:
        for b, L in enumerate(lens): s[b, :, L:] = -float("inf")
    want = torch.softmax(s, -1) @ v.double()
    worst = max(worst, (o.double() - want).abs().max().item())
B, T = 256, 512
sets = [tuple(torch.randn(B, T, 128, generator=g).cuda().to(torch.bfloat16) for _ in range(3)) for _ in range(3)]
for i in range(6): eng.attention(*sets[i %% 3])
torch.cuda.synchronize()
best = 1e9; res = []
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(30): eng.attention(*sets[i %% 3])
    e1.record(); torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / 30 * 1e3)
q, k, v = sets[0]
o = eng.attention(q, k, v).float()
want = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float())
big = (o - want).abs().max().item()
print("RESULT variant=%%s maxerr=%%.3e maxerr_bench_shape=%%.3e us=%%s" %% (os.environ.get("VADB_ATTN_VARIANT"), worst, big, " ".join("%%.1f" %% r for r in res)))
''' % root
variants = [int(a) for a in sys.argv[1:]] or [0, 1, 3, 5, 7, 9, 16, 18, 20, 22, 33, 35, 37, 39, 65, 69]
for v in variants:
    env = dict(os.environ, VADB_ATTN_VARIANT=str(v))
    try:
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=180, env=env)
        out = [l for l in (r.stdout + r.stderr).splitlines() if "RESULT" in l or "rror" in l or "timeout" in l]
        print(f"variant {v}: rc={r.returncode} ::", " | ".join(out[:4]), flush=True)
    except subprocess.TimeoutExpired:
        print(f"variant {v}: TIMEOUT", flush=True)
