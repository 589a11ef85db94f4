"""Secondary measurements (BASELINE.json configs 4 and 5, and the reference Predictor's window regime)."""
import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import vad_oracle as O
from vad_b200.engine import VadEngine

def timed(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

st = O.make_state(0, 64, 3, 128)
out = {}
for dtype in ("bf16", "fp32"):
    eng = VadEngine.from_state_dict(st, compute_dtype=dtype)
    g = torch.Generator().manual_seed(0)
    # config 4: long context T=8192, batch 64
    B, T = (64, 8192) if dtype == "bf16" else (8, 8192)
    x = (torch.randn(B, T, 64, generator=g) * 2 - 3).cuda()
    ms = timed(lambda: eng.forward(x, want_logp=False), iters=3)
    out[f"config4_T8192_B{B}_{dtype}"] = {"ms": ms, "frames_per_s": B * T / ms * 1e3,
                                          "attn_TFLOPs_equiv": 3 * 4 * B * T * T * 128 / ms / 1e9}
    # config 5: mixed lengths
    rnd = random.Random(0)
    lens = [rnd.choice([128, 512, 2048]) for _ in range(64)]
    Tm = max(lens)
    xm = (torch.randn(64, Tm, 64, generator=g) * 2 - 3).cuda()
    ln = torch.tensor(lens, dtype=torch.int32).cuda()
    ms = timed(lambda: eng.forward(xm, ln, want_logp=False), iters=3)
    valid = sum(lens)
    # parity over valid frames for a subset of clips (padded rows excluded), vs the oracle with the same mask
    sub = [0, 1, 2, 3]
    want = O.forward_prob(st, xm[sub].cpu(), torch.tensor([lens[i] for i in sub])).numpy()
    got = eng.forward(xm[sub].contiguous(), ln[sub].contiguous(), want_logp=False)[0].cpu().numpy()
    err = max(np.abs(got[i, :lens[s]] - want[i, :lens[s]]).max() for i, s in enumerate(sub))
    out[f"config5_mixed_B64_{dtype}"] = {"ms": ms, "valid_frames_per_s": valid / ms * 1e3, "padded_T": Tm,
                                         "max_abs_dP_valid": float(err)}
    eng.close()

# reference Predictor regime: 10 minutes of audio -> L = 60001 frames, F = 80, windows of 7
st80 = O.make_state(3, 80, 3, 128)
feat = O.make_input(5, 1, 60001, 80)[0].numpy()
for dtype in ("bf16", "fp32"):
    eng = VadEngine.from_state_dict(st80, compute_dtype=dtype)
    eng.predict_probabilities(feat, 19, 9)
    t0 = time.perf_counter(); n = 3
    for _ in range(n): probs, mean = eng.predict_probabilities(feat, 19, 9)
    dt = (time.perf_counter() - t0) / n
    out[f"window_path_L60001_{dtype}"] = {"s_per_call_host_to_host": dt, "audio_seconds_per_s": 600.0 / dt,
                                          "window_frames_per_s": (60001 - 38) * 7 / dt}
    eng.close()
torch.set_num_threads(16)
small = feat[:3038]                      # 30 s of audio for the CPU oracle
t0 = time.perf_counter(); ref = O.predict_probabilities(st80, small, 19, 9); dt = time.perf_counter() - t0
out["window_path_cpu_oracle_L3038"] = {"s": dt, "audio_seconds_per_s": 30.0 / dt}
eng = VadEngine.from_state_dict(st80, compute_dtype="bf16")
got, _ = eng.predict_probabilities(small, 19, 9)
out["window_path_parity_bf16_L3038"] = float(np.abs(got - ref).max())
import json; print(json.dumps(out, indent=1))
