"""Log-mel feature extraction -- the step UPSTREAM of the accelerated path ("next" row of
SURVEY.md section 8f).  Reference: vad/acoustics/feature_extractor.py:71-80 and
vad/acoustics/transforms/log_mel_spectrogram.py:19-32, i.e.
``np.log(librosa.feature.melspectrogram(y, sr, n_mels, n_fft, hop_length, win_length) + 1e-6)``
with librosa 0.8.0 defaults (hann window zero-padded to n_fft, center=True with reflect padding,
power=2, Slaney mel scale with Slaney area normalisation, fmin=0, fmax=sr/2).

librosa is not installed here and is not vendored by the reference, so this NumPy restatement of
its published algorithm has no golden vectors: **parity unpinned** for this step (the north-star
parity bar is stated "on identical log-mel inputs").  This module is the host-side (NumPy) FeatureExtractor
of the reference API; the predictor itself runs the same transform on the device
(csrc/k_logmel.cu via VadEngine.predict_audio / VadEngine.logmel) whenever n_fft is a power of two.
"""
from __future__ import annotations

import numpy as np

from .checkpoint import Config
from .data_models import AudioData


def _hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr: int, n_fft: int, n_mels: int) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels, fmin=0, fmax=sr/2, htk=False, norm='slaney')."""
    fftfreqs = np.linspace(0, sr / 2.0, 1 + n_fft // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(sr / 2.0), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    weights = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    return (weights * enorm[:, None]).astype(np.float32)


def log_mel_spectrogram(audio: np.ndarray, sr: int, n_fft: int, hop: int, win: int, n_mels: int):
    """-> [n_mels, frames] float32."""
    from scipy.signal import get_window
    window = get_window("hann", win, fftbins=True)
    lpad = (n_fft - win) // 2
    window = np.pad(window, (lpad, n_fft - win - lpad))
    y = np.pad(np.asarray(audio, dtype=np.float32), n_fft // 2, mode="reflect")
    n_frames = 1 + (len(y) - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n_frames)[:, None]
    spec = np.fft.rfft(y[idx] * window[None, :], axis=1).astype(np.complex64)
    power = (np.abs(spec) ** 2).T                         # [1+n_fft/2, frames]
    mel = mel_filterbank(sr, n_fft, n_mels).dot(power)
    return np.log(mel + 1e-6).astype(np.float32)


class LogMelSpectrogramTransform:
    def __init__(self, n_fft, hop_ms, window_ms, n_mels):
        self.n_fft, self.hop_ms, self.window_ms, self.n_mels = n_fft, hop_ms, window_ms, n_mels
        self.feature_size = n_mels

    def apply(self, audio_data: AudioData) -> np.ndarray:
        hop = int(self.hop_ms / 1000 * audio_data.sample_rate)
        win = int(self.window_ms / 1000 * audio_data.sample_rate)
        return log_mel_spectrogram(audio_data.audio, audio_data.sample_rate, self.n_fft, hop, win,
                                   self.n_mels)


class FeatureExtractor:
    """Inference subset of vad.acoustics.feature_extractor.FeatureExtractor: the log-mel
    transform without silence removal / SpecAugment / temporal differences (none of which the
    reference's inference checkpoints enable; anything else raises)."""

    def __init__(self, config, use_spec_augment: bool = False):
        self.config = Config.wrap(config)
        tr = self.config.get("transform")
        if tr is None or tr.get("name") != "log-mel":
            raise NotImplementedError("only the 'log-mel' transform is supported on this path")
        if self.config.get("temporal_differences") or self.config.get("silence_remover"):
            raise NotImplementedError("temporal differences / silence removal are not supported")
        self.transform = LogMelSpectrogramTransform(tr["n_fft"], tr["hop_ms"], tr["window_ms"],
                                                    tr["n_mels"])
        self.feature_size, self.feature_depth = self.transform.feature_size, 1

    def extract(self, audio_data: AudioData) -> np.ndarray:
        return self.transform.apply(audio_data)

    def extract_with_postprocessing(self, audio_data: AudioData) -> np.ndarray:
        # (feature_size, time) -> (time, feature_size): feature_extractor.py:77-80
        return np.ascontiguousarray(np.swapaxes(self.extract(audio_data), 0, 1))
