"""End-to-end smoke tests of the reference-facing surface on the GPU: from_checkpoint on a file in
the reference's checkpoint format, predict_probabilities / predict on synthetic audio, and the typer
CLI (`main.py predict`, `main.py evaluate`) -- the equivalents of the reference's tests/test_predict.py
and tests/test_evaluate.py."""
import json
import wave
from datetime import timedelta

import numpy as np
import pytest
import torch

from oracle import vad_oracle as O

pytestmark = pytest.mark.gpu

CFG = {"context_resolution": {"context_window_half_frames": 19, "context_window_jump_frames": 9,
                              "context_window_shift_frames": 39},
       "feature_extractor": {"silence_remover": None, "transform": {
           "name": "log-mel", "n_fft": 512, "hop_ms": 10, "window_ms": 25, "n_mels": 80, "n_mfcc": None},
           "spec_augment": None, "temporal_differences": False, "stack_differences": False, "cachedir": None},
       "model": {"name": "self-attention", "dnn": None, "boosted_dnn": None, "acam": None,
                 "self_attention": {"num_layers": 3, "d_model": 128, "dropout": 0.5}}}


def _write_wav(path, seconds=3.0, seed=0):
    rng = np.random.default_rng(seed)
    t = np.arange(int(16000 * seconds)) / 16000
    sig = 0.3 * np.sin(2 * np.pi * 220 * t) * (np.sin(2 * np.pi * 1.5 * t) > 0) + 0.01 * rng.standard_normal(len(t))
    pcm = (np.clip(sig, -1, 1) * 32767).astype(np.int16)
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000); w.writeframes(pcm.tobytes())


@pytest.fixture(scope="module")
def ckpt(tmp_path_factory):
    d = tmp_path_factory.mktemp("ck")
    st = O.make_state(3, 80, 3, 128)
    torch.save({"state_dict": st, "epoch": 1, "global_step": 2, "config": CFG,
                "metrics": {"val_auc": np.float64(0.5)}}, d / "sample.checkpoint")
    return d / "sample.checkpoint", st


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_predictor_from_checkpoint_matches_oracle(ckpt, tmp_path, dtype):
    from vad.predictor import VADFromScratchPredictor, VADPredictParameters     # reference import path
    from vad_b200.data_models import AudioData
    path, st = ckpt
    pred = VADFromScratchPredictor.from_checkpoint(path, torch.device("cuda"), compute_dtype=dtype)
    assert pred.context_window_frames == 7 and pred.feature_extractor.feature_size == 80
    wav = tmp_path / "a.wav"
    _write_wav(wav)
    audio = AudioData.load(wav)
    feat = pred.feature_extractor.extract_with_postprocessing(audio)
    probs = pred.predict_probabilities(audio)
    want = O.predict_probabilities(st, feat, 19, 9)
    assert probs.shape == want.shape == (len(feat), 7)
    assert np.abs(probs - want).max() <= (1e-3 if dtype == "fp32" else 1e-2)
    # model(features=...) keeps the reference call convention and returns log-probs [B,T,2]
    x = O.make_input(5, 3, 7, 80)
    logp = pred.model(features=x.cuda())
    assert logp.shape == (3, 7, 2)
    assert np.abs(logp.cpu().numpy() - O.forward_logp(st, x).numpy()).max() <= (2e-3 if dtype == "fp32" else 5e-2)
    va = pred.predict(audio, VADPredictParameters(None, 0.5, 0, 0, 0, 0, None, True, 100, False))
    assert va.duration == audio.duration and va.probs_sample_rate == 100 and len(va.probs) > 0


def test_cli_predict_and_evaluate(ckpt, tmp_path):
    from typer.testing import CliRunner
    from main import app
    from vad_b200.data_models import Activity, VoiceActivity
    path, _ = ckpt
    wav = tmp_path / "clip.wav"
    _write_wav(wav, 4.0, seed=1)
    out = tmp_path / "va.json"
    r = CliRunner().invoke(app, ["predict", str(wav), str(path), "--output-path", str(out), "--threshold", "0.3",
                                 "--return-probs", "--probs-sample-rate", "100"])
    assert r.exit_code == 0, r.output
    va = VoiceActivity.load(out)
    assert json.load(open(out))["version"] == "v0.3" and va.probs is not None
    # evaluate: one labelled pair
    truth = VoiceActivity(timedelta(seconds=4.0), [Activity(timedelta(seconds=0.5), timedelta(seconds=1.5)),
                                                   Activity(timedelta(seconds=2.2), timedelta(seconds=3.1))], None, None)
    truth.save(tmp_path / "clip.json")
    (tmp_path / "list.jsonl").write_text(json.dumps({"audio_path": "clip.wav", "voice_activity_path": "clip.json"}) + "\n")
    res = tmp_path / "eval.json"
    r = CliRunner().invoke(app, ["evaluate", str(tmp_path / "list.jsonl"), str(path), "--output-path", str(res)])
    assert r.exit_code == 0, r.output
    total = json.loads(open(res).readline())
    assert 0.0 <= total["auc"] <= 1.0 and "boosted_eer" in total


# ----------------------------------------------------------------------------------------------
# The reference's own end-to-end fixtures (its tests/test_predict.py:12-30): the real wav and the real
# test checkpoint's weights, goldens from the unmodified reference model (tests/golden/make_wav_golden.py)
# ----------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def real_ckpt(tmp_path_factory):
    from tests.golden_util import sample_checkpoint_state, weather_wav
    _, cfg, _ = weather_wav()
    d = tmp_path_factory.mktemp("realck")
    torch.save({"state_dict": sample_checkpoint_state(), "epoch": 0, "global_step": 2, "config": cfg,
                "metrics": {"val_auc": np.float64(0.585)}}, d / "sample.checkpoint")
    return d / "sample.checkpoint"


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_reference_wav_predict_probabilities_vs_reference_golden(real_ckpt, dtype):
    from tests.golden_util import weather_wav
    from vad.predictor import VADFromScratchPredictor
    audio, _, want = weather_wav()
    pred = VADFromScratchPredictor.from_checkpoint(real_ckpt, torch.device("cuda"), compute_dtype=dtype)
    probs = pred.predict_probabilities(audio)             # PCM -> log-mel -> windows -> model -> boost, on the device
    assert probs.shape == want.shape == (1022, 7)
    err = np.abs(probs - want).max()
    print(f"reference wav, {dtype}: max|dP| = {err:.3e}")
    assert err <= (1e-3 if dtype == "fp32" else 1e-2)
    np.testing.assert_array_equal(probs[want == 0.5], 0.5)


def test_reference_wav_cli_predict(real_ckpt, tmp_path):
    """Mirror of the reference's tests/test_predict.py: exit code 0 and at least one activity."""
    from typer.testing import CliRunner
    from main import app
    from tests.golden_util import WEATHER_WAV
    from vad_b200.data_models import VoiceActivity
    out = tmp_path / "va.json"
    r = CliRunner().invoke(app, ["predict", WEATHER_WAV, str(real_ckpt), "--output-path", str(out)])
    assert r.exit_code == 0, r.output
    va = VoiceActivity.load(out)
    assert len(va.activities) > 0
