"""Evaluation statistics used by ``main.py evaluate`` (vad/metrics.py:16-138).  CPU scalar
post-hoc statistics, outside the accelerated path; restated so the CLI is self-contained."""
from statistics import harmonic_mean

import numpy as np


def equal_error_rate(y_true, y_score):
    from scipy.interpolate import interp1d
    from scipy.optimize import brentq
    from sklearn.metrics import roc_curve
    fpr, tpr, _ = roc_curve(y_true, y_score, pos_label=1)
    return brentq(lambda x: 1 - x - interp1d(fpr, tpr)(x), 0, 1)


def detect_boundaries(frames):
    frames = np.asarray(frames).astype(np.int64)
    b = np.append(frames, 0) - np.append(0, frames)
    starts = np.where(b == 1)[0]
    ends = np.where(b == -1)[0] - 1
    return starts, ends, len(starts)


def _boundary_accuracy(true, pred, boundaries, num_segments, L, is_start):
    n = len(true)
    total = 0.0
    for bnd in boundaries:
        lo, hi = max(bnd - L, 0), min(bnd + L, n)
        idx = np.arange(lo, hi)
        w = ((idx - bnd) >= 0) if is_start else ((bnd - idx) >= 0)
        w = w.astype(np.float64)
        delta = (np.asarray(pred)[lo:hi] == np.asarray(true)[lo:hi]).astype(np.float64)
        total += (w * delta).sum() / w.sum()
    return total / num_segments if num_segments > 0 else 0


def vad_accuracy(frames_true, frames_pred, L=5):
    from sklearn.metrics import accuracy_score
    acc = accuracy_score(frames_true, frames_pred)
    starts, ends, n_true = detect_boundaries(frames_true)
    _, _, n_pred = detect_boundaries(frames_pred)
    sba = _boundary_accuracy(frames_true, frames_pred, starts, n_true, L, True)
    eba = _boundary_accuracy(frames_true, frames_pred, ends, n_true, L, False)
    bp = n_true / (2 * n_pred) * (sba + eba) if n_pred > 0 else 0
    vacc = harmonic_mean([acc, sba, eba, bp])
    return vacc, acc, sba, eba, bp
