"""CPU-side checks: the C-ABI library loads and exports every symbol include/vadb200.h declares
(no compute calls without a GPU), and the host-side logic (checkpoint reading, model parameter
naming, packing, sharding, post-processing, data models)."""
import ctypes
import json
import os
import re
from datetime import timedelta

import numpy as np
import pytest
import torch

from oracle import vad_oracle as O
from tests.golden_util import sample_checkpoint_state

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    from vad_b200 import _cabi
    if not os.path.exists(_cabi.LIB_PATH):
        g.build()
    return _cabi.load_library()


def test_library_exports_every_declared_symbol(lib):
    from vad_b200 import _cabi
    header = open(os.path.join(ROOT, "include", "vadb200.h")).read()
    declared = set(re.findall(r"\b(vadb_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_cabi.SYMBOLS), declared ^ set(_cabi.SYMBOLS)
    for sym in declared:
        assert getattr(lib, sym) is not None
    assert b"sm_100a" in lib.vadb_version()


def test_weight_count_matches_reference_state_dict(lib):
    from vad_b200 import _cabi
    from vad_b200.engine import pack_state
    cfg = _cabi.VadbConfig(80, 3, 128, _cabi.VADB_F32)
    n = lib.vadb_weight_count(ctypes.byref(cfg))
    st = sample_checkpoint_state()
    assert n == sum(v.numel() for v in st.values()) == 605698
    assert pack_state(st, 3).numel() == n
    cfg = _cabi.VadbConfig(64, 3, 128, _cabi.VADB_BF16)
    assert lib.vadb_weight_count(ctypes.byref(cfg)) == 603650


def test_create_fails_loudly_without_gpu(lib):
    from vad_b200 import _cabi
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    cfg = _cabi.VadbConfig(64, 3, 128, _cabi.VADB_BF16)
    rc = lib.vadb_create(ctypes.byref(h), ctypes.byref(cfg), 0)
    assert rc != 0 and not h.value
    assert b"CUDA" in lib.vadb_last_error(None)
    from vad_b200.engine import VadEngine
    with pytest.raises(RuntimeError):
        VadEngine(64, 3, 128)


def test_bad_config_rejected(lib):
    from vad_b200 import _cabi
    h = ctypes.c_void_p()
    cfg = _cabi.VadbConfig(64, 3, 256, _cabi.VADB_BF16)
    assert lib.vadb_create(ctypes.byref(h), ctypes.byref(cfg), 0) == -1
    assert b"d_model" in lib.vadb_last_error(None)


def test_model_parameter_names_match_reference_checkpoint():
    from vad_b200.model import SelfAttentiveVAD, create_model
    st = sample_checkpoint_state()
    m = SelfAttentiveVAD(80, 3, 128, 0.5)
    assert list(m.state_dict().keys()) == list(st.keys()) == O.state_keys(3)
    m.load_state_dict(st)                       # strict
    for k, v in m.state_dict().items():
        assert torch.equal(v, st[k])
    m2 = create_model({"name": "self-attention", "self_attention": {"num_layers": 2, "d_model": 128,
                                                                    "dropout": 0.1}}, 64, 7)
    assert m2.num_layers == 2 and m2.feature_size == 64
    with pytest.raises(NotImplementedError):
        create_model({"name": "acam"}, 64, 7)


def test_checkpoint_roundtrip_in_reference_format(tmp_path):
    """A file written in the reference's checkpoint layout (model_checkpointer.py:97-110)."""
    from vad_b200.checkpoint import Config, context_window_frames, load_checkpoint
    st = O.make_state(3, 80, 3, 128)
    cfg = {"context_resolution": {"context_window_half_frames": 19, "context_window_jump_frames": 9},
           "feature_extractor": {"silence_remover": None, "transform": {
               "name": "log-mel", "n_fft": 512, "hop_ms": 10, "window_ms": 25, "n_mels": 80, "n_mfcc": None},
               "temporal_differences": False, "stack_differences": False, "cachedir": None},
           "model": {"name": "self-attention", "self_attention": {"num_layers": 3, "d_model": 128, "dropout": 0.5}}}
    path = tmp_path / "x.checkpoint"
    torch.save({"state_dict": st, "epoch": 1, "global_step": 2, "config": cfg,
                "metrics": {"val_auc": np.float64(0.5)}}, path)
    ck = load_checkpoint(path)
    c = Config.wrap(ck["config"])
    assert c.model.self_attention.num_layers == 3 and c.feature_extractor.transform.n_mels == 80
    assert context_window_frames(c.context_resolution.context_window_half_frames,
                                 c.context_resolution.context_window_jump_frames) == 7
    assert torch.equal(ck["state_dict"]["classifier.weight"], st["classifier.weight"])


@pytest.mark.skipif(not os.path.exists("/root/reference/tests/checkpoints/vad/sample.checkpoint"),
                    reason="reference fixture not present on this box")
def test_reads_the_reference_sample_checkpoint():
    from vad_b200.checkpoint import load_checkpoint
    ck = load_checkpoint("/root/reference/tests/checkpoints/vad/sample.checkpoint")
    st = sample_checkpoint_state()
    assert list(ck["state_dict"].keys()) == list(st.keys())
    for k in st:
        assert torch.equal(ck["state_dict"][k], st[k])
    assert ck["config"]["model"]["name"] == "self-attention"


def test_shard_bounds_and_balance():
    from vad_b200.distributed import balanced_assignment, shard_bounds
    for n, w in [(2048, 8), (7, 4), (3, 8), (256, 1)]:
        spans = [shard_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    lengths = [128, 512, 2048] * 8
    parts = balanced_assignment(lengths, 4)
    assert sorted(i for p in parts for i in p) == list(range(24))
    loads = [sum(lengths[i] ** 2 for i in p) for p in parts]
    assert max(loads) / min(loads) < 1.1


def test_voice_activity_json_v03_roundtrip(tmp_path):
    from vad_b200.data_models import Activity, VoiceActivity
    va = VoiceActivity(timedelta(seconds=5.25), [Activity(timedelta(seconds=0.5), timedelta(seconds=1.75))],
                       None, None)
    p = tmp_path / "va.json"
    va.save(p)
    d = json.load(open(p))
    assert d["version"] == "v0.3" and d["duration"] == "00:00:05.250"
    assert d["activities"] == [{"start": "00:00:00.500", "end": "00:00:01.750"}]
    back = VoiceActivity.load(p)
    assert back.activities[0].end == timedelta(seconds=1.75)
    assert back.to_labels(100).sum() == 125


def _ref_pp():
    import importlib.util
    mods = {}
    for name in ("trim", "convert", "split"):
        path = f"/root/reference/vad/postprocessing/{name}.py"
        if not os.path.exists(path):
            pytest.skip("reference not present on this box")
        spec = importlib.util.spec_from_file_location("ref_" + name, path)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        mods[name] = m
    return mods


def test_postprocessing_bit_identical_to_reference_functions():
    import vad_b200.postprocessing as P
    ref = _ref_pp()
    rng = np.random.default_rng(0)
    for trial in range(120):
        n = int(rng.integers(1, 200))
        p = rng.random(n)
        if n > 5:
            w = int(rng.integers(1, 9))
            p = np.convolve(p, np.ones(w) / w, mode="same")
        pred = p > 0.5
        args = [int(rng.integers(0, 12)) for _ in range(4)]
        np.testing.assert_array_equal(ref["trim"].trim_voice_activity(pred, *args),
                                      P.trim_voice_activity(pred, *args))
        for hop, win in [(10, 25), (10, 10), (12.5, 25)]:
            np.testing.assert_array_equal(ref["convert"].convert_frames_to_samples(pred, 16000, hop, win),
                                          P.convert_frames_to_samples(pred, 16000, hop, win))
            pf = p.astype(np.float32)
            np.testing.assert_array_equal(ref["convert"].convert_frames_to_samples(pf, 100, hop, win),
                                          P.convert_frames_to_samples(pf, 100, hop, win))
        sp = ref["convert"].convert_frames_to_samples(pred, 1000, 10, 25)
        assert ref["convert"].convert_samples_to_segments(sp, 1000) == P.convert_samples_to_segments(sp, 1000)
        ps = ref["convert"].convert_frames_to_samples(p, 1000, 10, 25)
        for mx, sr in ((1, 50), (1, 200), (2, 100), (1, 7)):
            np.testing.assert_array_equal(ref["split"].optimal_split_voice_activity(sp, ps, mx, sr),
                                          P.optimal_split_voice_activity(sp, ps, mx, sr))


def test_postprocessing_known_answers():
    """Hand-checked cases of the reference semantics (run on the GPU box too)."""
    import vad_b200.postprocessing as P
    x = np.array([0, 1, 1, 0, 0, 1, 0, 0, 0, 0, 1, 1, 1, 0], dtype=bool)
    np.testing.assert_array_equal(P.trim_voice_activity(x, 3, 0, 0, 0).astype(int),
                                  [0, 1, 1, 1, 1, 1, 0, 0, 0, 0, 1, 1, 1, 0])
    np.testing.assert_array_equal(P.trim_voice_activity(x, 0, 2, 0, 0).astype(int),
                                  [0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 0])
    # hang_over alone does nothing (the reference's `hang_before > 0 or hang_before > 0`)
    np.testing.assert_array_equal(P.trim_voice_activity(x, 0, 0, 0, 2), x)
    np.testing.assert_array_equal(P.trim_voice_activity(x, 0, 0, 1, 1).astype(int),
                                  [1, 1, 1, 1, 1, 1, 1, 0, 0, 1, 1, 1, 1, 1])
    s = P.convert_frames_to_samples(np.array([1, 0, 1], dtype=bool), 1000, 10, 25)
    assert len(s) == 45 and s[0] == 1 and s[12] == 0.5 and s[22] == pytest.approx(2 / 3) and s[44] == 1
    segs = P.convert_samples_to_segments(np.array([0, 1, 1, 0.5, 0, 0, 1, 1.0]), 10)
    assert segs == [(timedelta(seconds=0.1), timedelta(seconds=0.3)),
                    (timedelta(seconds=0.6), timedelta(seconds=0.7))]


def test_log_mel_shapes_and_filterbank():
    from vad_b200.data_models import AudioData
    from vad_b200.features import FeatureExtractor, mel_filterbank
    fb = mel_filterbank(16000, 512, 80)
    assert fb.shape == (80, 257) and (fb >= 0).all() and (fb.sum(axis=1) > 0).all()
    fe = FeatureExtractor({"transform": {"name": "log-mel", "n_fft": 512, "hop_ms": 10, "window_ms": 25,
                                         "n_mels": 80, "n_mfcc": None},
                           "temporal_differences": False, "stack_differences": False})
    audio = np.sin(2 * np.pi * 440 * np.arange(80000) / 16000).astype(np.float32)
    feat = fe.extract_with_postprocessing(AudioData(audio, 16000, timedelta(seconds=5)))
    assert feat.shape == (501, 80) and feat.dtype == np.float32     # 5 s -> 501 frames
    assert np.isfinite(feat).all() and feat.min() >= np.log(1e-6) - 1e-3
    assert abs(int(feat.mean(axis=0).argmax()) - 11) <= 1          # 440 Hz lands in mel band ~11


def test_logmel_host_tables_match_oracle(lib):
    """The mel filterbank and window the CUDA log-mel kernel uses are built on the host by the library
    (csrc/k_logmel.cu::logmel_tables); they must be the NumPy restatement's, bit for bit."""
    from oracle import logmel_oracle as LO
    for sr, n_fft, win, n_mels in [(16000, 512, 400, 80), (16000, 512, 400, 64), (8000, 256, 200, 40),
                                   (16000, 1024, 400, 128), (44100, 2048, 1102, 96)]:
        fb = np.empty((n_mels, n_fft // 2 + 1), np.float32)
        w = np.empty(n_fft, np.float64)
        rc = lib.vadb_logmel_tables(sr, n_fft, win, n_mels, fb.ctypes.data_as(ctypes.c_void_p),
                                    w.ctypes.data_as(ctypes.c_void_p))
        assert rc == 0
        np.testing.assert_array_equal(fb, LO.mel_filterbank(sr, n_fft, n_mels))
        np.testing.assert_allclose(w, LO.padded_window(n_fft, win), rtol=0, atol=1e-15)
    assert lib.vadb_logmel_tables(16000, 500, 400, 80, None, None) != 0          # n_fft not a power of two
    assert lib.vadb_logmel_frames(80000, 160) == 501 and lib.vadb_logmel_frames(0, 160) == 1


def test_logmel_oracle_matches_host_feature_extractor():
    """oracle/logmel_oracle.py and the product's host-side FeatureExtractor restate the same algorithm."""
    from oracle import logmel_oracle as LO
    from vad_b200.features import log_mel_spectrogram
    a = (np.random.default_rng(0).standard_normal(16000 * 2) * 0.1).astype(np.float32)
    np.testing.assert_array_equal(LO.log_mel_frames(a, 16000, 512, 160, 400, 80),
                                  log_mel_spectrogram(a, 16000, 512, 160, 400, 80).T)


def test_product_synthetic_factories_match_the_oracle():
    """bench.py's GPU arm draws its random weights / inputs from the product package
    (vad_b200.synthetic), its cpu_baseline leg from the oracle: both must be the same model and data."""
    from vad_b200 import synthetic as S
    for F in (64, 80):
        a, b = O.make_state(0, F, 3, 128), S.random_state(0, F, 3, 128)
        assert list(a.keys()) == list(b.keys())
        assert all(torch.equal(a[k], b[k]) for k in a)
    assert torch.equal(O.make_input(1, 2, 16, 64), S.random_features(1, 2, 16, 64))


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the GPU arm) needs no GPU: one JSON
    line with the contract's keys; under torchrun only rank 0 prints."""
    import subprocess
    import sys
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0
    assert d["metric"] == "audio frames/sec (T=512,F=64)" and d["higher_is_better"] is True
    # the unmodified reference modules when oracle/_ref travelled with the snapshot, else the oracle port
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "vad", "models", "self_attention.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    # a non-zero rank exits quietly
    r1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1"],
                        capture_output=True, text=True, timeout=600, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r1.returncode == 0 and r1.stdout.strip() == ""


def test_oracle_ref_recipe_and_agreement_with_port():
    """oracle/_ref (the reference's own model files, placed by oracle/build_ref.py; git-ignored) is what
    `bench.py --impl reference` times when present: it must be the reference class, not the repository's
    compatibility shim, and agree with the oracle port on the bench's sample."""
    import subprocess
    import sys
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "vad", "models", "self_attention.py")):
        pytest.skip("oracle/_ref not built (no /root/reference at build time)")
    code = (
        "import sys, torch; sys.path.insert(0, %r)\n"
        "from oracle.build_ref import import_reference_model\n"
        "from oracle import vad_oracle as O\n"
        "M = import_reference_model(); assert M is not None and issubclass(M, torch.nn.Module)\n"
        "st = O.make_state(0, 64, 3, 128); m = M(64, 3, 128, 0.5); m.load_state_dict(st); m.eval()\n"
        "x = O.make_input(1, 2, 128, 64)\n"
        "with torch.no_grad(): got = m(features=x)\n"
        "print(float((got - O.forward_logp(st, x)).abs().max()))\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert float(r.stdout.strip().splitlines()[-1]) <= 2e-6
