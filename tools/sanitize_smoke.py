"""Small end-to-end exercise of every device path (forward, ragged forward, window/audio path) for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from vad_b200 import synthetic as S
from vad_b200.engine import VadEngine
eng = VadEngine.from_state_dict(S.random_state(0, 64, 3, 128), compute_dtype="bf16")
g = torch.Generator().manual_seed(0)
x = (torch.randn(300, 256, 64, generator=g)).cuda().to(torch.bfloat16)       # 600 tiles: > 148 CTAs x 4
p, lp = eng.forward(x)
torch.cuda.synchronize()
# ragged, odd shapes, host lengths
x2 = torch.randn(9, 300, 64, generator=g).cuda()
p2, _ = eng.forward(x2, np.asarray([300, 1, 128, 129, 256, 257, 77, 5, 300]), want_logp=False)
torch.cuda.synchronize()
# window path + audio path
a = np.random.default_rng(0).standard_normal(16000 * 3).astype(np.float32) * 0.1
eng80 = VadEngine.from_state_dict(S.random_state(1, 80, 3, 128), compute_dtype="bf16")
probs = eng80.predict_audio(a, 16000, 512, 160, 400, 19, 9)[0]
print("ok", float(p.float().mean()), float(p2.float().nanmean()), probs.shape)
