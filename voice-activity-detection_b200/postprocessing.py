"""Frame -> segment post-processing DOWNSTREAM of the accelerated path ("next" row of SURVEY.md
section 8f): vectorised (run-length based) NumPy restatements of the reference's sequential
Python scans, bit-identical in result:

  trim_voice_activity          vad/postprocessing/trim.py:4-66
  convert_frames_to_samples    vad/postprocessing/convert.py:6-24
  convert_samples_to_segments  vad/postprocessing/convert.py:27-61
  optimal_split_voice_activity vad/postprocessing/split.py:26-104

The reference iterates every audio *sample* in Python (16 k iterations per second of audio);
these run in a handful of array passes and stay on the CPU (they are scalar bookkeeping).
"""
from __future__ import annotations

from datetime import timedelta

import numpy as np


def _edges(x: np.ndarray):
    """Indices i >= 1 of rising (x[i-1]==0, x[i]==1) and falling (x[i-1]==1, x[i]==0) edges,
    using the reference's exact ``== 0`` / ``== 1`` tests."""
    if len(x) < 2:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    prev, cur = x[:-1], x[1:]
    rising = np.nonzero((prev == 0) & (cur == 1))[0] + 1
    falling = np.nonzero((prev == 1) & (cur == 0))[0] + 1
    return rising, falling


def trim_voice_activity(predictions, min_vally=20, min_hill=20, hang_before=10, hang_over=10):
    out = predictions.copy()
    n = len(out)

    if min_vally > 0:      # fill valleys shorter than min_vally that lie between two hills
        rising, falling = _edges(out)
        if len(rising) and len(falling):
            # for each rising edge, the closest falling edge before it
            j = np.searchsorted(falling, rising, side="left") - 1
            ok = j >= 0
            starts, ends = falling[j[ok]], rising[ok]
            for s, e in zip(starts[(ends - starts) < min_vally], ends[(ends - starts) < min_vally]):
                out[s:e] = 1

    if min_hill > 0:       # flatten hills shorter than min_hill that lie between two valleys
        rising, falling = _edges(out)
        if len(rising) and len(falling):
            j = np.searchsorted(rising, falling, side="left") - 1
            ok = j >= 0
            starts, ends = rising[j[ok]], falling[ok]
            keep = (ends - starts) < min_hill
            for s, e in zip(starts[keep], ends[keep]):
                out[s:e] = 0

    # the reference tests ``hang_before > 0 or hang_before > 0`` (trim.py:46): hang_over alone
    # never triggers the pass -- kept as is
    if hang_before > 0:
        rising, falling = _edges(out)      # edges of the snapshot; writes do not feed back
        for i in rising:
            out[(0 if i < hang_before else i - hang_before):i] = 1
        for i in falling:
            if n - hang_over < i:
                out[i:] = 1
            else:
                out[i:i + hang_over] = 1
    return out


def convert_frames_to_samples(frames, sample_rate=16000, hop_ms=10, window_ms=10):
    frames = np.asarray(frames)
    hop = sample_rate * hop_ms / 1000
    win = sample_rate * window_ms / 1000
    nf = len(frames)
    num_samples = int((nf - 1) * hop + win)
    samples = np.zeros(num_samples)
    counts = np.zeros(num_samples)
    if nf == 0 or num_samples <= 0:
        return samples
    # window starts follow the reference's running float accumulation (start_index += hop)
    starts_f = np.concatenate([[0.0], np.cumsum(np.full(nf - 1, hop, dtype=np.float64))]) \
        if float(hop).is_integer() else _running_sum(hop, nf)
    starts = starts_f.astype(np.int64)
    ends = np.minimum((starts_f + win).astype(np.int64), num_samples)
    # every sample is covered by a run of consecutive frames; add them in frame order so the
    # float64 sums round exactly as the reference's sequential "+=" does
    s_idx = np.arange(num_samples)
    first = np.searchsorted(ends, s_idx, side="right")           # first frame with end > s
    last = np.searchsorted(starts, s_idx, side="right") - 1      # last frame with start <= s
    depth = int((last - first).max()) + 1 if num_samples else 0
    vals = frames.astype(np.float64)
    for k in range(depth):
        j = first + k
        m = j <= last
        samples[m] = samples[m] + vals[j[m]]
        counts[m] += 1
    counts[counts == 0] = 1
    return samples / counts


def _running_sum(step, n):
    out = np.empty(n, dtype=np.float64)
    acc = 0.0
    for i in range(n):
        out[i] = acc
        acc += step
    return out


def _voice_runs(samples):
    """State machine of convert.py:38-56 / split.py:42-52: a run starts at the first sample equal
    to 1 while not in voice and ends at the first later sample equal to 0; samples that are
    neither 0 nor 1 change nothing.  Returns (start_idx, end_idx) with end_idx = index of the
    terminating 0, or -1 for a run still open at the end."""
    samples = np.asarray(samples)
    ev = np.zeros(len(samples), dtype=np.int8)
    ev[samples == 1] = 1
    ev[samples == 0] = -1
    pos = np.nonzero(ev)[0]
    if len(pos) == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    e = ev[pos]
    keep = np.ones(len(e), dtype=bool)
    keep[1:] = e[1:] != e[:-1]                # collapse repeats: only state changes matter
    pos, e = pos[keep], e[keep]
    if e[0] == -1:                            # leading zeros while not in voice: no-op
        pos, e = pos[1:], e[1:]
    starts = pos[0::2]
    ends = pos[1::2]
    if len(ends) < len(starts):
        ends = np.concatenate([ends, [-1]])
    return starts.astype(np.int64), ends.astype(np.int64)


def convert_samples_to_segments(samples, sample_rate=16000):
    starts, ends = _voice_runs(samples)
    segments = []
    last = len(samples) - 1
    for s, e in zip(starts, ends):
        start_time = timedelta(seconds=int(s) / sample_rate)
        end_time = timedelta(seconds=(int(e) - 1) / sample_rate) if e >= 0 \
            else timedelta(seconds=last / sample_rate)
        segments.append((start_time, end_time))
    return segments


def optimal_split_long_block(block_sample_probs, max_samples):
    """split.py:80-104 with an explicit stack instead of recursion."""
    assert max_samples > 1
    half = max_samples // 2
    out = []
    stack = [(0, len(block_sample_probs))]
    while stack:
        lo, hi = stack.pop()
        block = block_sample_probs[lo:hi]
        bp = half + int(np.argmin(block[half:-half]))
        out.append(lo + bp)
        if bp > max_samples:                          # left block = block[:bp]
            stack.append((lo, lo + bp))
        if (hi - lo) - (bp + 1) > max_samples:        # right block = block[bp+1:]
            stack.append((lo + bp + 1, hi))
    return sorted(out)


def optimal_split_voice_activity(sample_predictions, sample_probs, max_length_seconds=300,
                                 sample_rate=16000):
    max_samples = max_length_seconds * sample_rate
    split_predictions = sample_predictions.copy()
    n = min(len(sample_predictions), len(sample_probs))       # zip() semantics
    starts, ends = _voice_runs(np.asarray(sample_predictions)[:n])
    for s, e in zip(starts, ends):
        e = n if e < 0 else int(e)
        if e - s > max_samples:
            for bp in optimal_split_long_block(sample_probs[s:e], max_samples):
                split_predictions[s + bp] = 0
    return split_predictions
