"""The oracle restatement vs. outputs of the real reference modules (tests/golden)."""
import numpy as np
import pytest
import torch

from oracle import vad_oracle as O
from tests.golden_util import golden, model_cases, predictor_cases, sample_checkpoint_state, valid_mask

CASES = model_cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_forward_matches_reference(name):
    if name == "syn_t8192":
        torch.set_num_threads(8)
    mk_state, mk_x, len_key = CASES[name]
    g = golden()
    lengths = g[len_key] if len_key else None
    got = O.forward_logp(mk_state(), mk_x(), lengths).numpy()
    want = g[name]
    assert got.shape == want.shape
    m = valid_mask(want.shape[:2], lengths)
    # same ATen kernels, same op order -> agreement to fp32 rounding
    np.testing.assert_allclose(got[m], want[m], rtol=0, atol=2e-6)
    if lengths is not None:
        # padded query rows also match (they attend to the valid keys; nothing is NaN)
        assert np.isfinite(got).all()
        np.testing.assert_allclose(got, want, rtol=0, atol=2e-6)


@pytest.mark.parametrize("name", sorted(predictor_cases()))
def test_predict_probabilities_matches_reference(name):
    feat = predictor_cases()[name]()
    got = O.predict_probabilities(sample_checkpoint_state(), feat, 19, 9)
    want = golden()[name]
    assert got.shape == want.shape == (len(feat), 7)
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-6)


def test_window_geometry():
    # vad/predictor.py:57-59 with half=19, jump=9 -> W=7, rel=[-19,-10,-1,0,1,10,19]
    assert O.context_window_frames(19, 9) == 7
    assert O.relative_neighbors(19, 9).tolist() == [-19, -10, -1, 0, 1, 10, 19]
    feat = np.arange(50 * 3, dtype=np.float32).reshape(50, 3)
    w, p = O.gather_windows(feat, 19, 9)
    assert w.shape == (12, 7, 3) and p.shape == (12, 7)
    assert p[0].tolist() == [0, 9, 18, 19, 20, 29, 38]
    assert p[-1].tolist() == [11, 20, 29, 30, 31, 40, 49]
    np.testing.assert_array_equal(w[3, 2], feat[p[3, 2]])


def test_boost_unfilled_slots_are_half():
    # slots never written stay (0,0) -> softmax -> 0.5 (vad/predictor.py:239-258)
    L, W = 60, 7
    feat = np.zeros((L, 4), np.float32)
    _, pos = O.gather_windows(feat, 19, 9)
    outputs = np.log(np.full((len(pos), W, 2), [0.25, 0.75], dtype=np.float32))
    b = O.boosted_aggregate(outputs, pos, L, W)
    filled = np.zeros((L, W), bool)
    filled[pos, np.arange(W)[None, :].repeat(len(pos), 0)] = True
    np.testing.assert_allclose(b[filled], 0.75, atol=1e-6)
    np.testing.assert_allclose(b[~filled], 0.5, atol=0)
    assert (~filled).any() and filled.any()


def test_positional_encoding_values():
    pe = O.positional_encoding(5, 128)[0]
    assert pe.shape == (5, 128)
    assert torch.allclose(pe[0, 0::2], torch.zeros(64)) and torch.allclose(pe[0, 1::2], torch.ones(64))
    assert abs(float(pe[3, 0]) - np.sin(3.0)) < 1e-6 and abs(float(pe[3, 1]) - np.cos(3.0)) < 1e-6


def test_mask_from_lengths():
    m = O.mask_from_lengths(torch.tensor([1, 3]), 4)
    assert m.tolist() == [[False, True, True, True], [False, False, False, True]]


def test_reference_wav_fixture_matches_reference():
    """The reference's own end-to-end fixture (its tests/test_predict.py wav + test checkpoint): the oracle's
    predictor on this repository's host log-mel features against the UNMODIFIED reference model run through
    the literal transcription of vad/predictor.py:169-258 (tests/golden/make_wav_golden.py)."""
    from tests.golden_util import weather_wav
    audio, cfg, want = weather_wav()
    from vad_b200.features import FeatureExtractor
    feat = FeatureExtractor(cfg["feature_extractor"]).extract_with_postprocessing(audio)
    assert feat.shape == (1022, 80)
    got = O.predict_probabilities(sample_checkpoint_state(), feat, 19, 9)
    assert got.shape == want.shape == (1022, 7)
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-6)
