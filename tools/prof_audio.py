"""(developer tool; uses the test oracle's NumPy log-mel only as the host-side comparison)
Time the whole reference-predictor path from PCM: audio -> log-mel -> windows -> model -> boosted
probabilities (vad/predictor.py:159-262), device front end vs the host (NumPy) feature extraction."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from vad_b200 import synthetic as S
from oracle import logmel_oracle as LO
from vad_b200.engine import VadEngine

seconds = float(os.environ.get("SECONDS_AUDIO", 600))
st = S.random_state(3, 80, 3, 128)
eng = VadEngine.from_state_dict(st, compute_dtype="bf16")
a = (np.random.default_rng(0).standard_normal(int(16000 * seconds)) * 0.1).astype(np.float32)
for _ in range(2):
    eng.predict_audio(a, 16000, 512, 160, 400, 19, 9)
t0 = time.perf_counter(); n = 5
for _ in range(n):
    probs, mean, _ = eng.predict_audio(a, 16000, 512, 160, 400, 19, 9)
dev_ms = (time.perf_counter() - t0) / n * 1e3
ad = torch.from_numpy(a).cuda()
for _ in range(2):
    eng.logmel(ad, 16000, 512, 160, 400, 80)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    eng.logmel(ad, 16000, 512, 160, 400, 80)
e1.record(); torch.cuda.synchronize()
lm_ms = e0.elapsed_time(e1) / 10
t0 = time.perf_counter()
feat = LO.log_mel_frames(a, 16000, 512, 160, 400, 80)
host_feat_ms = (time.perf_counter() - t0) * 1e3
t0 = time.perf_counter()
eng.predict_probabilities(feat, 19, 9)
host_rest_ms = (time.perf_counter() - t0) * 1e3
print(f"audio path, {seconds:.0f} s of 16 kHz PCM ({len(a) * 4 / 1e6:.1f} MB): device front end {dev_ms:.2f} ms host to host "
      f"({seconds / dev_ms * 1e3:.0f}x real time; log-mel kernel alone {lm_ms:.3f} ms for {feat.shape[0]} frames); "
      f"NumPy features {host_feat_ms:.0f} ms + model {host_rest_ms:.2f} ms")
