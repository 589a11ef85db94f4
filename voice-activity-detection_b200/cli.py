"""typer commands ``predict`` and ``evaluate`` with the reference's options
(vad/predict.py:10-50, vad/evaluate.py:20-185)."""
import json
import random
from collections import OrderedDict
from pathlib import Path
from typing import Optional

import numpy as np
import torch
from typer import Option

from .data_models import AudioData, VADDataList, VoiceActivity
from .metrics import equal_error_rate, vad_accuracy
from .predictor import VADFromScratchPredictor, VADPredictParameters


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("vad_b200 needs a CUDA device (B200); there is no CPU fallback")
    return torch.device("cuda")


def predict_vad_from_scratch(
    audio_path: Path,
    checkpoint_path: Path,
    output_path: Optional[Path] = Option(None, help="Path to store output. Default to stdout."),
    split_max_seconds: Optional[float] = Option(None, help="Chunk size to split audio in seconds."),
    activity_max_sec: Optional[int] = Option(None, help="Maximum length of voice activity in seconds"),
    threshold: float = 0.5,
    min_vally_ms: int = 0,
    min_hill_ms: int = 0,
    hang_before_ms: int = 0,
    hang_over_ms: int = 0,
    return_probs: bool = False,
    probs_sample_rate: Optional[int] = None,
    compute_dtype: str = Option("bf16", help="bf16 (tensor cores) or fp32 (parity path)"),
):
    predictor = VADFromScratchPredictor.from_checkpoint(checkpoint_path, _device(), compute_dtype)
    voice_activity = predictor.predict_from_path(
        audio_path,
        VADPredictParameters(split_max_seconds, threshold, min_vally_ms, min_hill_ms, hang_before_ms,
                             hang_over_ms, activity_max_sec, return_probs, probs_sample_rate, True))
    if output_path:
        output_path.parent.mkdir(parents=True, exist_ok=True)
        voice_activity.save(path=output_path)
    else:
        print(voice_activity)


_KEYS = ["auc", "accuracy", "precision", "recall", "vacc", "sba", "eba", "bp", "eer"]


def evaluate_vad_from_scratch(
    eval_path: Path,
    checkpoint_path: Path,
    output_path: Optional[Path] = Option(None, help="Path to store output. Default to stdout."),
    data_dir: Optional[Path] = None,
    threshold: float = 0.5,
    shuffle: bool = False,
    limit: Optional[int] = None,
    random_seed: int = 0,
    compute_dtype: str = Option("bf16", help="bf16 (tensor cores) or fp32 (parity path)"),
):
    from sklearn.metrics import accuracy_score, precision_score, recall_score, roc_auc_score
    predictor = VADFromScratchPredictor.from_checkpoint(checkpoint_path, _device(), compute_dtype)
    if data_dir is None:
        data_dir = eval_path.parent
    pairs = VADDataList.load(eval_path).pairs
    if shuffle:
        random.seed(random_seed)
        random.shuffle(pairs)
    if limit:
        pairs = pairs[:limit]
    results = []
    for pair in pairs:
        audio_path = data_dir.joinpath(pair.audio_path)
        va_path = data_dir.joinpath(pair.voice_activity_path)
        true_labels = VoiceActivity.load(va_path).to_labels(100)
        probs = predictor.predict_probabilities(AudioData.load(audio_path))
        single = probs[:, int(probs.shape[1] / 2)][: len(true_labels)]        # evaluate.py:57-59
        single_pred = single > threshold
        boosted = probs.mean(axis=1)[: len(true_labels)]                      # evaluate.py:61-62
        boosted_pred = boosted > threshold
        true_labels = true_labels[: len(boosted)]
        r = OrderedDict(audio_path=str(audio_path), voice_activity_path=str(va_path))
        # the reference computes the un-boosted block from the boosted arrays too (evaluate.py:65-68)
        r["auc"] = roc_auc_score(true_labels, boosted)
        r["accuracy"] = accuracy_score(true_labels, boosted_pred)
        r["precision"] = precision_score(true_labels, boosted_pred)
        r["recall"] = recall_score(true_labels, boosted_pred)
        r["vacc"], _, r["sba"], r["eba"], r["bp"] = vad_accuracy(true_labels, single_pred)
        r["eer"] = equal_error_rate(true_labels, single_pred)
        r["boosted_auc"], r["boosted_accuracy"] = r["auc"], r["accuracy"]
        r["boosted_precision"], r["boosted_recall"] = r["precision"], r["recall"]
        (r["boosted_vacc"], _, r["boosted_sba"], r["boosted_eba"],
         r["boosted_bp"]) = vad_accuracy(true_labels, boosted_pred)
        r["boosted_eer"] = equal_error_rate(true_labels, boosted_pred)
        print(f"\n{pair.audio_path}")
        for k in _KEYS:
            print(f"{k.upper() if len(k) <= 4 else k.capitalize()}: {r[k]:0.2%}")
        for k in _KEYS:
            print(f"Boosted {k.upper() if len(k) <= 4 else k.capitalize()}: {r['boosted_' + k]:0.2%}")
        results.append(r)
    total = {k: float(np.mean([r[k] for r in results])) for k in
             _KEYS + ["boosted_" + k for k in _KEYS]}
    print("\nTotal:")
    for k, v in total.items():
        print(f"{k}: {v:0.2%}")
    if output_path is not None:
        output_path.parent.mkdir(parents=True, exist_ok=True)
        with output_path.open("w") as f:
            f.write(json.dumps(total, ensure_ascii=False) + "\n")
            for r in results:
                f.write(json.dumps(r, ensure_ascii=False) + "\n")
