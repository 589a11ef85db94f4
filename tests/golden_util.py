"""Shared helpers for the golden fixtures (tests/golden/*.npz, made by make_golden.py)."""
import os

import numpy as np
import torch

from oracle import vad_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden():
    return np.load(os.path.join(GOLDEN_DIR, "reference_outputs.npz"))


def sample_checkpoint_state():
    z = np.load(os.path.join(GOLDEN_DIR, "sample_checkpoint_state.npz"))
    return {k: torch.from_numpy(z[k]) for k in O.state_keys(3)}


# name -> (state factory, input factory, lengths key or None); mirrors make_golden.py
def model_cases():
    syn = lambda: O.make_state(0, 64, 3, 128)
    sharp = lambda: O.make_state(21, 64, 3, 128, ln_jitter=0.3, weight_gain=4.0)
    return {
        "ckpt_w7": (sample_checkpoint_state, lambda: O.make_input(11, 64, 7, 80), None),
        "ckpt_t512": (sample_checkpoint_state, lambda: O.make_input(12, 2, 512, 80), None),
        "syn_t512": (syn, lambda: O.make_input(1, 4, 512, 64), None),
        "syn_t128": (syn, lambda: O.make_input(2, 3, 128, 64), None),
        "syn_t1": (syn, lambda: O.make_input(3, 2, 1, 64), None),
        "syn_t300": (syn, lambda: O.make_input(4, 2, 300, 64), None),
        "syn_t2048": (syn, lambda: O.make_input(5, 1, 2048, 64), None),
        "syn_t8192": (syn, lambda: O.make_input(6, 1, 8192, 64), None),
        "syn_masked": (syn, lambda: O.make_input(7, 4, 512, 64), "syn_masked_lengths"),
        "sharp_t512": (sharp, lambda: O.make_input(8, 2, 512, 64), None),
        "sharp_masked": (sharp, lambda: O.make_input(9, 3, 384, 64), "sharp_masked_lengths"),
        "l2_f80_t64": (lambda: O.make_state(31, 80, 2, 128),
                       lambda: O.make_input(10, 5, 64, 80), None),
    }


def predictor_cases():
    return {
        "ckpt_predict_probs": lambda: O.make_input(13, 1, 101, 80)[0].numpy(),
        "ckpt_predict_probs_long": lambda: O.make_input(14, 1, 1203, 80)[0].numpy(),
        "ckpt_predict_probs_short": lambda: O.make_input(15, 1, 30, 80)[0].numpy(),
    }


def valid_mask(shape_BT, lengths):
    B, T = shape_BT
    if lengths is None:
        return np.ones((B, T), dtype=bool)
    return np.arange(T)[None, :] < np.asarray(lengths)[:, None]


def prob_from_logp(logp):
    e = np.exp(logp - logp.max(axis=-1, keepdims=True))
    return (e / e.sum(axis=-1, keepdims=True))[..., 1]


WEATHER_WAV = os.path.join(GOLDEN_DIR, "When_the_Weather_Is_Fine_12_4.wav")


def weather_wav():
    """The reference's tests/test_predict.py fixture: (AudioData, checkpoint config dict, golden [L,7])."""
    import json
    from vad_b200.data_models import AudioData
    z = np.load(os.path.join(GOLDEN_DIR, "weather_wav_golden.npz"))
    return AudioData.load(WEATHER_WAV), json.loads(str(z["config_json"])), z["probs"]
