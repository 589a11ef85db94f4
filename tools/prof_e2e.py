"""Where does the host-to-host call spend its time?  (H2D bandwidth, chunk count sweep, small-batch forwards)"""
import os, sys, time, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vad_b200 import synthetic as S

def ev_time(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

def wall(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(iters): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e3

if len(sys.argv) > 1 and sys.argv[1] == "child":
    from vad_b200.engine import VadEngine
    st = S.random_state(0, 64, 3, 128)
    eng = VadEngine.from_state_dict(st, compute_dtype="bf16")
    xs = [S.random_features(i, 256, 512, 64).pin_memory() for i in range(3)]
    k = [0]
    def f():
        k[0] += 1
        eng.forward(xs[k[0] % 3], want_logp=False)
    print(json.dumps({"chunks": os.environ.get("VADB_HOST_CHUNKS", "default"), "ms": wall(f, 20, 5)}))
    sys.exit(0)

out = {}
h = torch.empty(256, 512, 64).pin_memory(); d = torch.empty_like(h, device="cuda")
out["h2d_33MB_ms"] = ev_time(lambda: d.copy_(h, non_blocking=True))
hp = torch.empty(256, 512).pin_memory(); dp = torch.empty_like(hp, device="cuda")
out["d2h_0.5MB_ms"] = ev_time(lambda: hp.copy_(dp, non_blocking=True))
from vad_b200.engine import VadEngine
st = S.random_state(0, 64, 3, 128)
eng = VadEngine.from_state_dict(st, compute_dtype="bf16")
for B in (32, 37, 64, 74, 86, 111, 128, 148, 256):
    x = S.random_features(1, B, 512, 64).cuda()
    out[f"dev_forward_B{B}_ms"] = ev_time(lambda: eng.forward(x, want_logp=False))
    out[f"dev_forward_B{B}_wall_ms"] = wall(lambda: eng.forward(x, want_logp=False))
eng.close()
print(json.dumps(out, indent=1))
for c in ("default", "1", "2", "3", "4", "5", "6", "8"):
    env = dict(os.environ)
    if c != "default": env["VADB_HOST_CHUNKS"] = c
    print(subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True).stdout.strip())
