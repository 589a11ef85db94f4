"""Generate golden vectors by running the REAL reference modules (unmodified, imported from
/root/reference) on seeded inputs.  Run once in the build container:

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box, so its outputs are committed as small
fixtures next to this script.  Inputs and synthetic weights are regenerated from seeds by
``oracle.vad_oracle.make_input`` / ``make_state`` (the generator loads those exact weights
into the reference ``SelfAttentiveVAD`` via ``load_state_dict``), so the fixtures hold only
outputs (+ the reference test checkpoint's state_dict, its one real-weights fixture).
"""
import _codecs
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("VAD_REFERENCE", "/root/reference")
sys.path.insert(0, REF)                     # reference's own `vad` package first
from vad.modeling.transformer import mask_from_lengths  # noqa: E402  (reference)
from vad.models.self_attention import SelfAttentiveVAD  # noqa: E402  (reference)

sys.path.insert(1, ROOT)
from oracle import vad_oracle as O  # noqa: E402

torch.set_num_threads(8)


def ref_model(state):
    F_, L, d = O.infer_dims(state)
    m = SelfAttentiveVAD(F_, L, d, 0.5)     # = vad/models/model_factory.py:42-48
    m.load_state_dict(state)
    return m.eval()


def ref_forward(model, x, lengths=None):
    with torch.no_grad():
        if lengths is None:
            return model(features=x)        # keyword as in vad/predictor.py:224
        mask = mask_from_lengths(torch.as_tensor(lengths), max_length=x.shape[1])
        h = model.input_layer(x)
        h = model.encoder(h, sources_key_padding_mask=mask)
        return model.log_softmax(model.classifier(h))


def load_sample_checkpoint():
    with torch.serialization.safe_globals([
            (np._core.multiarray.scalar, "numpy.core.multiarray.scalar"), np.dtype,
            _codecs.encode, type(np.dtype("float64"))]):
        return torch.load(os.path.join(REF, "tests/checkpoints/vad/sample.checkpoint"),
                          map_location="cpu")


def predictor_loop_transcription(model, feature, half, jump):
    """Per-item restatement of vad/predictor.py:169-258 (the reference file itself cannot be
    imported here: omegaconf / more_itertools / librosa are absent).  Kept in the
    reference's item-by-item form on purpose, to pin the vectorised oracle."""
    L = len(feature)
    W = 2 * (half - 1) // jump + 3
    n = L - 2 * half
    outs, poss = [], []
    for start in range(0, max(n, 0), 1000):
        items = range(start, min(start + 1000, n))
        feats, pos = [], []
        for item in items:
            center = half + item
            rel = np.concatenate([np.arange(-half, 0, jump), np.array([0]),
                                  np.arange(1, half + 1, jump)], axis=0)
            feats.append(feature[center + rel])
            pos.append(center + rel)
        batch = torch.from_numpy(np.stack(feats))
        with torch.no_grad():
            out = model(features=batch)
        outs.append(out.numpy())
        poss.append(np.stack(pos))
    boosted = np.zeros((L, W, 2), dtype=np.float32)
    for out, pos in zip(outs, poss):
        widx = np.expand_dims(np.arange(W), 0).repeat(len(pos), axis=0)
        boosted[pos, widx] = out
    from scipy.special import softmax
    return softmax(boosted, axis=2)[:, :, 1]


def main():
    out = {}
    meta = {}

    # (1) reference test checkpoint (F=80), Predictor-shaped call [n,7,80] and [2,512,80]
    ck = load_sample_checkpoint()
    sd = {k: v.clone() for k, v in ck["state_dict"].items()}
    np.savez(os.path.join(HERE, "sample_checkpoint_state.npz"),
             **{k: v.numpy() for k, v in sd.items()})
    m = ref_model(sd)
    x = O.make_input(11, 64, 7, 80)
    out["ckpt_w7"] = ref_forward(m, x).numpy()
    x = O.make_input(12, 2, 512, 80)
    out["ckpt_t512"] = ref_forward(m, x).numpy()
    feat = O.make_input(13, 1, 101, 80)[0].numpy()          # a 1 s "clip": 101 frames
    out["ckpt_predict_probs"] = predictor_loop_transcription(m, feat, 19, 9)
    feat = O.make_input(14, 1, 1203, 80)[0].numpy()         # >1 chunk of 1000 windows
    out["ckpt_predict_probs_long"] = predictor_loop_transcription(m, feat, 19, 9)
    feat = O.make_input(15, 1, 30, 80)[0].numpy()           # L < 2*half: no windows at all
    out["ckpt_predict_probs_short"] = predictor_loop_transcription(m, feat, 19, 9)

    # (2) seeded synthetic weights, BASELINE shapes (F=64, d=128, L=3)
    st = O.make_state(0, 64, 3, 128)
    m = ref_model(st)
    out["syn_t512"] = ref_forward(m, O.make_input(1, 4, 512, 64)).numpy()
    out["syn_t128"] = ref_forward(m, O.make_input(2, 3, 128, 64)).numpy()
    out["syn_t1"] = ref_forward(m, O.make_input(3, 2, 1, 64)).numpy()
    out["syn_t300"] = ref_forward(m, O.make_input(4, 2, 300, 64)).numpy()   # ragged tile
    out["syn_t2048"] = ref_forward(m, O.make_input(5, 1, 2048, 64)).numpy()
    out["syn_t8192"] = ref_forward(m, O.make_input(6, 1, 8192, 64)).numpy()
    # (3) masked mixed-length batch (config 5), by-hand call chain
    lengths = [128, 512, 300, 1]
    out["syn_masked"] = ref_forward(m, O.make_input(7, 4, 512, 64), lengths).numpy()
    meta["syn_masked_lengths"] = np.array(lengths)
    # (4) sharper softmax: larger attention weights force online-softmax rescaling paths
    st2 = O.make_state(21, 64, 3, 128, ln_jitter=0.3, weight_gain=4.0)
    m2 = ref_model(st2)
    out["sharp_t512"] = ref_forward(m2, O.make_input(8, 2, 512, 64)).numpy()
    out["sharp_masked"] = ref_forward(m2, O.make_input(9, 3, 384, 64), [384, 200, 129]).numpy()
    meta["sharp_masked_lengths"] = np.array([384, 200, 129])
    # (5) other model sizes accepted by the config (num_layers, d_model fixed by kernels = 128)
    st3 = O.make_state(31, 80, 2, 128)
    out["l2_f80_t64"] = ref_forward(ref_model(st3), O.make_input(10, 5, 64, 80)).numpy()

    np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **out, **meta)
    for k, v in out.items():
        print(k, v.shape, float(np.abs(v).max()))


if __name__ == "__main__":
    main()
