"""VADFromScratchPredictor: the reference's predictor API (vad/predictor.py:27-304) with the hot
loop -- window gather, batched model forward, boosted aggregation -- executed on the B200 by
libvadb200 in ONE device call per audio chunk (no per-frame Python loop, no per-1000-window
host<->device round trips).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from datetime import timedelta
from itertools import chain
from pathlib import Path
from typing import List, Optional, Union

import numpy as np
import torch

from .checkpoint import Config, context_window_frames, load_checkpoint
from .data_models import Activity, AudioData, VoiceActivity
from .features import FeatureExtractor
from .model import SelfAttentiveVAD, create_model
from .postprocessing import (convert_frames_to_samples, convert_samples_to_segments,
                             optimal_split_voice_activity, trim_voice_activity)


@dataclass
class VADPredictParameters:            # field order is positional in vad/predict.py:32-43
    split_max_seconds: Optional[float]
    threshold: float
    min_vally_ms: int
    min_hill_ms: int
    hang_before_ms: int
    hang_over_ms: int
    activity_max_seconds: Optional[int]
    return_probs: bool
    probs_sample_rate: Optional[int]
    show_progress_bar: bool


class VADFromScratchPredictor:
    def __init__(self, model: SelfAttentiveVAD, feature_extractor: FeatureExtractor,
                 device: torch.device, config):
        self.model = model
        self.feature_extractor = feature_extractor
        self.device = device
        self.config = Config.wrap(config)
        cr = self.config.context_resolution
        self.context_window_half_frames = cr.context_window_half_frames
        self.context_window_jump_frames = cr.context_window_jump_frames
        self.context_window_frames = context_window_frames(self.context_window_half_frames,
                                                           self.context_window_jump_frames)
        # a transform is always configured on this path (vad/predictor.py:66-69)
        self.feature_window_half_size = self.context_window_half_frames
        self.feature_window_jump_size = self.context_window_jump_frames
        self.feature_window_one_unit = 1

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_checkpoint(cls, checkpoint_path: Path, device: torch.device,
                        compute_dtype: str = "bf16"):
        """vad/predictor.py:264-280.  ``device`` must be a CUDA device: there is no CPU path."""
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("vad_b200 predictor needs a CUDA device (B200); there is no CPU fallback")
        checkpoint = load_checkpoint(checkpoint_path, map_location="cpu")
        config = Config.wrap(checkpoint["config"])
        feature_extractor = FeatureExtractor(config.feature_extractor, use_spec_augment=False)
        W = context_window_frames(config.context_resolution.context_window_half_frames,
                                  config.context_resolution.context_window_jump_frames)
        model = create_model(config.model, feature_extractor.feature_size, W, compute_dtype)
        model.load_state_dict(checkpoint["state_dict"])
        model = model.to(device=device)
        return cls(model=model, feature_extractor=feature_extractor, device=device, config=config)

    # ------------------------------------------------------------------ hot path
    def predict_probabilities(self, audio_data: Union[AudioData, np.ndarray, torch.Tensor]) -> np.ndarray:
        """-> positive-class probabilities [L, W] float32 (vad/predictor.py:159-262).
        Accepts AudioData (features are extracted first, :160) or an already extracted feature
        matrix [L, F] (numpy or tensor)."""
        if self.config.model.name != "self-attention":
            raise NotImplementedError
        eng = self.model.engine(self.device)
        if isinstance(audio_data, AudioData):
            tr = self.feature_extractor.transform
            n_fft = int(tr.n_fft)
            if n_fft >= 32 and n_fft <= 4096 and (n_fft & (n_fft - 1)) == 0 and \
                    os.environ.get("VADB_HOST_FEATURES", "0") != "1":
                # audio -> log-mel -> windows -> model -> boosted probabilities, all on the device:
                # only the PCM samples go up and the [L, W] probabilities come back
                sr = audio_data.sample_rate
                probs, _, _ = eng.predict_audio(audio_data.audio, sr, n_fft, int(tr.hop_ms / 1000 * sr),
                                                int(tr.window_ms / 1000 * sr),
                                                self.context_window_half_frames,
                                                self.context_window_jump_frames)
                return probs
            feature = self.feature_extractor.extract_with_postprocessing(audio_data)
        else:
            feature = audio_data
        probs, _ = eng.predict_probabilities(feature, self.context_window_half_frames,
                                             self.context_window_jump_frames)
        if isinstance(probs, torch.Tensor):
            probs = probs.cpu().numpy()
        return probs

    # ------------------------------------------------------------------ full predict
    def predict_from_path(self, audio_path: Path, parameters: VADPredictParameters) -> VoiceActivity:
        return self.predict(AudioData.load(Path(audio_path)), parameters)

    def predict(self, audio_data: AudioData, parameters: VADPredictParameters) -> VoiceActivity:
        """vad/predictor.py:77-157 (chunking, boosting by the window mean, thresholding, trim,
        frame->sample->segment conversion, optional optimal split, optional probs)."""
        if parameters.split_max_seconds is not None:
            num_chunks = math.ceil(audio_data.duration.total_seconds() / parameters.split_max_seconds)
        else:
            num_chunks = 1
        num_chunks = max(num_chunks, 1)
        adjusted = audio_data.duration.total_seconds() / num_chunks
        tr = self.feature_extractor.config.transform
        hop_ms, window_ms = tr.hop_ms, tr.window_ms
        chunks = []
        for ci in range(num_chunks):
            s = int(ci * adjusted * audio_data.sample_rate)
            e = int((ci + 1) * adjusted * audio_data.sample_rate)
            chunk = AudioData(audio_data.audio[s:e], sample_rate=audio_data.sample_rate,
                              duration=timedelta(seconds=adjusted))
            frame_probabilities = self.predict_probabilities(chunk)
            boosted = frame_probabilities.mean(axis=1)
            predictions = boosted > parameters.threshold
            trimmed = trim_voice_activity(
                predictions,
                min_vally=round(parameters.min_vally_ms / hop_ms),
                min_hill=round(parameters.min_hill_ms / hop_ms),
                hang_before=round(parameters.hang_before_ms / hop_ms),
                hang_over=round(parameters.hang_over_ms / hop_ms))
            sample_predictions = convert_frames_to_samples(trimmed, sample_rate=16000, hop_ms=hop_ms,
                                                           window_ms=window_ms)
            if parameters.activity_max_seconds is not None and parameters.activity_max_seconds > 0:
                sample_full_probs = convert_frames_to_samples(boosted, sample_rate=16000,
                                                              hop_ms=hop_ms, window_ms=window_ms)
                sample_predictions = optimal_split_voice_activity(
                    sample_predictions=sample_predictions, sample_probs=sample_full_probs,
                    max_length_seconds=parameters.activity_max_seconds, sample_rate=16000)
            segments = convert_samples_to_segments(sample_predictions, sample_rate=16000)
            activities = [Activity(start=a, end=b) for a, b in segments]
            probs = None
            if parameters.return_probs:
                probs = convert_frames_to_samples(boosted, sample_rate=parameters.probs_sample_rate,
                                                  hop_ms=hop_ms, window_ms=window_ms).tolist()
            chunks.append(VoiceActivity(
                duration=chunk.duration, activities=activities,
                probs_sample_rate=parameters.probs_sample_rate if parameters.return_probs else None,
                probs=probs))
        return merge_voice_activities(chunks)


def merge_voice_activities(voice_activities: List[VoiceActivity]) -> VoiceActivity:
    """vad/predictor.py:283-304."""
    offset = timedelta(0)
    activities = []
    for va in voice_activities:
        activities += [Activity(start=a.start + offset, end=a.end + offset) for a in va.activities]
        offset += va.duration
    probs = None
    if voice_activities[0].probs:
        probs = list(chain(*[va.probs for va in voice_activities]))
    return VoiceActivity(duration=sum([va.duration for va in voice_activities], timedelta(0)),
                         activities=activities,
                         probs_sample_rate=voice_activities[0].probs_sample_rate, probs=probs)
