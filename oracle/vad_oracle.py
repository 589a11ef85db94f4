"""CPU oracle for the Self-Attentive VAD hot path.  TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement (torch fp32 on CPU for the model arithmetic, NumPy for
the index work) of the reference algorithm for the one hot path this repository
accelerates.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  The product path
(``vad_b200``) never does: it fails loudly when ``libvadb200.so`` is missing.

Parity pin: the restatement is checked against outputs of the *real* reference modules
(``/root/reference/vad/models/self_attention.py`` + ``vad/modeling/transformer.py``,
imported unmodified in the build container by ``tests/golden/make_golden.py``) stored as
fixtures under ``tests/golden/``; ``tests/test_oracle_golden.py`` replays them.  The
Predictor-level functions (window gather / boosted aggregation) cannot be imported from
the reference here (``vad/predictor.py`` needs omegaconf, more_itertools, librosa), so
they are restated from the source and pinned against a literal transcription of the
reference's per-item loop semantics in the golden generator.  Log-mel extraction
(librosa 0.8.0, absent) is outside the boundary: **parity unpinned** for that step.

Every function cites the reference file:line it follows (paths relative to the
reference repository root).

The op *sequence* of the reference is kept on purpose (three materialised [B,1,T,T]
score tensors, division by a float64 ``np.sqrt`` scalar, a ``.contiguous()`` copy ...)
so that timing this port on host cores is a faithful CPU baseline of the reference's
own PyTorch CPU forward.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

State = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------------------
# Parameter naming (vad/models/self_attention.py:12-21, vad/modeling/transformer.py:10-61,
# 227-254, 366-375).  Same keys as the reference ``state_dict``.
# --------------------------------------------------------------------------------------
def state_keys(num_layers: int):
    keys = ["input_layer.0.weight", "input_layer.0.bias"]
    for l in range(num_layers):
        p = f"encoder.layers.{l}."
        for proj in ("query", "key", "value", "final"):
            keys += [p + f"self_attention.{proj}_projection.weight",
                     p + f"self_attention.{proj}_projection.bias"]
        keys += [p + "self_attention_sublayer.layer_norm.weight",
                 p + "self_attention_sublayer.layer_norm.bias",
                 p + "feed_forward.feed_forward.0.weight",
                 p + "feed_forward.feed_forward.0.bias",
                 p + "feed_forward.feed_forward.3.weight",
                 p + "feed_forward.feed_forward.3.bias",
                 p + "feed_forward_sublayer.layer_norm.weight",
                 p + "feed_forward_sublayer.layer_norm.bias"]
    keys += ["encoder.layer_norm.weight", "encoder.layer_norm.bias",
             "classifier.weight", "classifier.bias"]
    return keys


def state_shapes(feature_size: int, num_layers: int, d_model: int):
    d, dff = d_model, 4 * d_model  # d_ff = 4*d_model: vad/models/self_attention.py:10
    shapes = OrderedDict()
    for k in state_keys(num_layers):
        if k == "input_layer.0.weight":
            shapes[k] = (d, feature_size)
        elif k == "classifier.weight":
            shapes[k] = (2, d)
        elif k == "classifier.bias":
            shapes[k] = (2,)
        elif k.endswith("feed_forward.0.weight"):
            shapes[k] = (dff, d)
        elif k.endswith("feed_forward.0.bias"):
            shapes[k] = (dff,)
        elif k.endswith("feed_forward.3.weight"):
            shapes[k] = (d, dff)
        elif k.endswith("projection.weight"):
            shapes[k] = (d, d)
        else:
            shapes[k] = (d,)
    return shapes


def make_state(seed: int, feature_size: int = 64, num_layers: int = 3, d_model: int = 128,
               ln_jitter: float = 0.1, weight_gain: float = 1.0) -> State:
    """Deterministic synthetic weights (nn.Linear-style U(-1/sqrt(in), 1/sqrt(in)) init,
    LayerNorm gamma/beta jittered so that the affine terms are exercised).  The golden
    generator loads exactly this state into the real reference module, so fixtures only
    need to store (seed, shape) and outputs."""
    g = torch.Generator().manual_seed(seed)
    st: State = OrderedDict()
    for k, shp in state_shapes(feature_size, num_layers, d_model).items():
        if "layer_norm.weight" in k:
            st[k] = 1.0 + ln_jitter * (2 * torch.rand(shp, generator=g) - 1)
        elif "layer_norm.bias" in k:
            st[k] = ln_jitter * (2 * torch.rand(shp, generator=g) - 1)
        else:
            fan_in = shp[-1] if len(shp) == 2 else state_shapes(feature_size, num_layers, d_model)[
                k.replace(".bias", ".weight")][-1]
            bound = weight_gain / math.sqrt(fan_in)
            st[k] = (2 * torch.rand(shp, generator=g) - 1) * bound
        st[k] = st[k].to(torch.float32).contiguous()
    return st


def make_input(seed: int, B: int, T: int, Fdim: int) -> torch.Tensor:
    """Synthetic log-mel-like input of SURVEY.md section 8(d): randn*2 - 3, fp32."""
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B, T, Fdim, generator=g) * 2.0 - 3.0).to(torch.float32)


def infer_dims(state: State) -> Tuple[int, int, int]:
    d_model, feature_size = state["input_layer.0.weight"].shape
    num_layers = 0
    while f"encoder.layers.{num_layers}.self_attention.query_projection.weight" in state:
        num_layers += 1
    return int(feature_size), int(num_layers), int(d_model)


# --------------------------------------------------------------------------------------
# Model forward
# --------------------------------------------------------------------------------------
def positional_encoding(length: int, d_model: int) -> torch.Tensor:
    """vad/modeling/transformer.py:403-414 (build_positional_encoding).  fp32, [1,length,d]."""
    pe = torch.zeros(length, d_model)
    position = torch.arange(0, length, dtype=torch.float32).unsqueeze(1)
    div_term = torch.exp(
        torch.arange(0, d_model, 2, dtype=torch.float32) * -(math.log(10000.0) / d_model)
    )
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0)


def mask_from_lengths(lengths: torch.Tensor, max_length: Optional[int] = None) -> torch.Tensor:
    """vad/modeling/transformer.py:432-447.  True = masked (j >= length[b])."""
    if max_length is None:
        max_length = int(lengths.max().item())
    positions = torch.arange(max_length).unsqueeze(0)
    return torch.ge(positions, lengths.unsqueeze(1))


def _attention(st: State, p: str, x: torch.Tensor, key_padding_mask) -> torch.Tensor:
    """MultiHeadAttention.forward with n_heads = 1 (vad/models/self_attention.py:17-19):
    vad/modeling/transformer.py:281-284 (projections), :305-314 (head views), :351-363
    (scaled_dot_product), :319-325 (key padding mask), :333 (softmax over keys), :338-347."""
    B, T, d = x.shape
    n_heads, d_head = 1, d
    q = F.linear(x, st[p + "query_projection.weight"], st[p + "query_projection.bias"])
    k = F.linear(x, st[p + "key_projection.weight"], st[p + "key_projection.bias"])
    v = F.linear(x, st[p + "value_projection.weight"], st[p + "value_projection.bias"])
    qh = q.view(B, T, n_heads, d_head).transpose(1, 2)
    kh = k.view(B, T, n_heads, d_head).transpose(1, 2)
    vh = v.view(B, T, n_heads, d_head).transpose(1, 2)
    dot = torch.matmul(qh, kh.transpose(2, 3))
    scores = dot / np.sqrt(d_head)                      # :362, float64 numpy scalar divisor
    if key_padding_mask is not None:
        m = key_padding_mask.unsqueeze(1).unsqueeze(1).expand_as(scores)
        scores = scores.masked_fill(m, float("-inf"))  # :319-325
    attn = torch.softmax(scores, dim=3)                 # :254, :333
    ctx = torch.matmul(attn, vh)                        # :338
    ctx = ctx.transpose(1, 2).contiguous().view(B, T, d)  # :341-346
    return F.linear(ctx, st[p + "final_projection.weight"], st[p + "final_projection.bias"])


def encoder_forward(st: State, h: torch.Tensor, key_padding_mask=None) -> torch.Tensor:
    """TransformerEncoder.forward (vad/modeling/transformer.py:24-34) over
    TransformerEncoderLayer.forward (:49-61) with Sublayer (:234-238): pre-LayerNorm,
    residual added to the UN-normalised value; dropout is identity in eval."""
    _, L, d = infer_dims(st)
    for l in range(L):
        p = f"encoder.layers.{l}."
        a = F.layer_norm(h, (d,), st[p + "self_attention_sublayer.layer_norm.weight"],
                         st[p + "self_attention_sublayer.layer_norm.bias"], 1e-5)
        h = _attention(st, p + "self_attention.", a, key_padding_mask) + h
        f = F.layer_norm(h, (d,), st[p + "feed_forward_sublayer.layer_norm.weight"],
                         st[p + "feed_forward_sublayer.layer_norm.bias"], 1e-5)
        # PositionwiseFeedForwardNetwork: vad/modeling/transformer.py:370-375
        f = F.linear(f, st[p + "feed_forward.feed_forward.0.weight"],
                     st[p + "feed_forward.feed_forward.0.bias"])
        f = torch.relu(f)
        f = F.linear(f, st[p + "feed_forward.feed_forward.3.weight"],
                     st[p + "feed_forward.feed_forward.3.bias"])
        h = f + h
    return F.layer_norm(h, (d,), st["encoder.layer_norm.weight"],
                        st["encoder.layer_norm.bias"], 1e-5)


def forward_logp(st: State, features: torch.Tensor,
                 lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
    """SelfAttentiveVAD.forward (vad/models/self_attention.py:23-28) -> log-probs [B,T,2].

    ``lengths`` (optional, [B] int) reproduces the masked call chain that SURVEY.md
    section 0 describes: input_layer -> encoder(x, sources_key_padding_mask=
    mask_from_lengths(lengths, T)) -> classifier -> log_softmax."""
    _, _, d = infer_dims(st)
    with torch.no_grad():
        x = features.to(torch.float32)
        T = x.shape[1]
        h = F.linear(x, st["input_layer.0.weight"], st["input_layer.0.bias"])  # :13
        h = h + positional_encoding(T, d)[:, :T] / math.sqrt(d)  # transformer.py:390,401
        mask = None
        if lengths is not None:
            mask = mask_from_lengths(torch.as_tensor(lengths, dtype=torch.int64), T)
        h = encoder_forward(st, h, mask)
        z = F.linear(h, st["classifier.weight"], st["classifier.bias"])
        return torch.log_softmax(z, dim=2)


def forward_prob(st: State, features: torch.Tensor,
                 lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
    """P(speech) as the callers take it: softmax(logp)[..., 1]
    (vad/predictor.py:225, :257-258)."""
    return torch.softmax(forward_logp(st, features, lengths), dim=-1)[..., 1]


# --------------------------------------------------------------------------------------
# Predictor-level index work (vad/predictor.py)
# --------------------------------------------------------------------------------------
def context_window_frames(half: int, jump: int) -> int:
    """vad/predictor.py:57-59 / :270-275."""
    return 2 * (half - 1) // jump + 3


def relative_neighbors(half: int, jump: int) -> np.ndarray:
    """vad/predictor.py:186-199 (feature) == :205-216 (label) when a transform is set
    (feature_window_one_unit == 1, :66-69): [-half..0) step jump, 0, [1..half+1) step jump."""
    left = np.arange(-half, 0, jump)
    right = np.arange(1, half + 1, jump)
    return np.concatenate([left, np.array([0]), right], axis=0)


def gather_windows(feature: np.ndarray, half: int, jump: int):
    """vad/predictor.py:169, :180-220: one window per centre frame half+i,
    i in [0, L-2*half).  Returns (windows [n,W,F] f32, positions [n,W] i64)."""
    L = len(feature)
    n = max(L - 2 * half, 0)
    rel = relative_neighbors(half, jump)
    centers = half + np.arange(n)
    positions = centers[:, None] + rel[None, :]
    windows = feature[positions] if n > 0 else np.zeros((0, len(rel), feature.shape[1]),
                                                        dtype=feature.dtype)
    return windows.astype(np.float32), positions.astype(np.int64)


def boosted_aggregate(outputs: np.ndarray, positions: np.ndarray, label_length: int,
                      W: int) -> np.ndarray:
    """vad/predictor.py:238-258: scatter log-probs into boosted_outputs[L,W,2] (zeros
    elsewhere), softmax over the class axis, take class 1 -> [L,W].  Unfilled slots stay
    (0,0) -> 0.5; ``boosted_counts`` is computed by the reference but never used."""
    boosted = np.zeros((label_length, W, 2), dtype=np.float32)
    if len(outputs):
        widx = np.arange(W)[None, :].repeat(len(positions), axis=0)
        boosted[positions, widx] = outputs
    m = boosted.max(axis=2, keepdims=True)
    e = np.exp(boosted - m)
    probs = e / e.sum(axis=2, keepdims=True)   # scipy.special.softmax(axis=2)
    return probs[:, :, 1].astype(np.float32)


def predict_probabilities(st: State, feature: np.ndarray, half: int, jump: int,
                          chunk_size: int = 1000) -> np.ndarray:
    """VADFromScratchPredictor.predict_probabilities (vad/predictor.py:159-262) from the
    log-mel feature matrix [L,F] on (i.e. after :160).  Returns [L, W] float32."""
    L = len(feature)
    W = context_window_frames(half, jump)
    windows, positions = gather_windows(np.asarray(feature, dtype=np.float32), half, jump)
    outs = []
    for s in range(0, len(windows), chunk_size):          # ichunked(range(n), 1000) :182
        outs.append(forward_logp(st, torch.from_numpy(windows[s:s + chunk_size])).numpy())
    outputs = np.concatenate(outs, axis=0) if outs else np.zeros((0, W, 2), np.float32)
    return boosted_aggregate(outputs, positions, L, W)


def boosted_mean(probs_LW: np.ndarray) -> np.ndarray:
    """vad/predictor.py:95 and vad/evaluate.py:61: mean over the window axis."""
    return probs_LW.mean(axis=1)
