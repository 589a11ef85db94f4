"""CPU oracle for the log-mel front end.  TEST INFRASTRUCTURE ONLY:
only ``tests/`` and ``__graft_entry__.smoke()`` may import it.

NumPy restatement of the reference's feature extraction for the log-mel transform:

  vad/acoustics/transforms/log_mel_spectrogram.py:19-32
      np.log(librosa.feature.melspectrogram(y, sr, n_mels, n_fft, hop_length, win_length) + 1e-6)
  vad/acoustics/feature_extractor.py:71-80   (feature_size, time) -> (time, feature_size)

with librosa 0.8.0's published defaults (stft: periodic hann window of win_length zero-padded and
centred to n_fft, center=True with reflect padding, numpy.fft.rfft in float64 cast to complex64;
melspectrogram: power=2; filters.mel: Slaney scale, Slaney area normalisation, fmin=0, fmax=sr/2).

**Parity unpinned**: librosa 0.8.0 is neither vendored by the reference nor installed in the build
container, and the reference's tests hold no feature fixtures, so this restatement could not be checked
against librosa itself.  What is pinned: the CUDA kernel (csrc/k_logmel.cu) against this file, and the
library's host-built filterbank / window tables against the functions below.
"""
from __future__ import annotations

import numpy as np


def hz_to_mel(f):
    """librosa.core.convert.hz_to_mel(htk=False)."""
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, f / f_sp)


def mel_to_hz(m):
    """librosa.core.convert.mel_to_hz(htk=False)."""
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr: int, n_fft: int, n_mels: int) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels, fmin=0, fmax=sr/2, htk=False, norm='slaney') -> [n_mels, 1+n_fft/2]."""
    fftfreqs = np.linspace(0, sr / 2.0, 1 + n_fft // 2)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(0.0), hz_to_mel(sr / 2.0), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    weights = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    return (weights * enorm[:, None]).astype(np.float32)


def padded_window(n_fft: int, win: int) -> np.ndarray:
    """scipy.signal.get_window('hann', win, fftbins=True) zero-padded (centred) to n_fft, float64."""
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(win) / win)
    lpad = (n_fft - win) // 2
    return np.pad(w, (lpad, n_fft - win - lpad))


def log_mel_frames(audio: np.ndarray, sr: int, n_fft: int, hop: int, win: int, n_mels: int) -> np.ndarray:
    """-> [frames, n_mels] float32 (the layout FeatureExtractor.extract_with_postprocessing returns)."""
    window = padded_window(n_fft, win)
    y = np.pad(np.asarray(audio, dtype=np.float32), n_fft // 2, mode="reflect")
    n_frames = 1 + (len(y) - n_fft) // hop
    out = np.empty((n_frames, n_mels), dtype=np.float32)
    fb = mel_filterbank(sr, n_fft, n_mels)
    step = 4096
    for f0 in range(0, n_frames, step):                      # chunked: bounded memory for long clips
        f1 = min(n_frames, f0 + step)
        idx = np.arange(n_fft)[None, :] + hop * np.arange(f0, f1)[:, None]
        spec = np.fft.rfft(y[idx] * window[None, :], axis=1).astype(np.complex64)
        power = (np.abs(spec) ** 2).T                        # [1+n_fft/2, frames] float32
        out[f0:f1] = np.log(fb.dot(power) + 1e-6).astype(np.float32).T
    return out
