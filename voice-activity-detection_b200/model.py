"""SelfAttentiveVAD: the reference model's interface (vad/models/self_attention.py:6-28,
vad/models/model_factory.py:10-54) over the B200 engine.

The module carries parameters under the reference's names so that reference checkpoints load
with ``load_state_dict`` unchanged, but its ``forward`` runs no torch arithmetic: it hands the
features to ``VadEngine`` (hand-written sm_100a kernels behind libvadb200's C ABI).
"""
from __future__ import annotations

from enum import Enum
from typing import Optional

import torch
from torch import Tensor, nn

from .engine import VadEngine


class _Attention(nn.Module):       # parameter holder for vad/modeling/transformer.py:241-254
    def __init__(self, d):
        super().__init__()
        self.query_projection = nn.Linear(d, d)
        self.key_projection = nn.Linear(d, d)
        self.value_projection = nn.Linear(d, d)
        self.final_projection = nn.Linear(d, d)


class _Sublayer(nn.Module):        # vad/modeling/transformer.py:227-232
    def __init__(self, d):
        super().__init__()
        self.layer_norm = nn.LayerNorm(d)


class _FeedForward(nn.Module):     # vad/modeling/transformer.py:366-375 (indices 0 and 3 hold weights)
    def __init__(self, d, dff, dropout):
        super().__init__()
        self.feed_forward = nn.Sequential(nn.Linear(d, dff), nn.ReLU(), nn.Dropout(dropout),
                                          nn.Linear(dff, d))


class _EncoderLayer(nn.Module):    # vad/modeling/transformer.py:37-47
    def __init__(self, d, dff, dropout):
        super().__init__()
        self.self_attention = _Attention(d)
        self.self_attention_sublayer = _Sublayer(d)
        self.feed_forward = _FeedForward(d, dff, dropout)
        self.feed_forward_sublayer = _Sublayer(d)


class _Encoder(nn.Module):         # vad/modeling/transformer.py:10-22
    def __init__(self, num_layers, d, dff, dropout):
        super().__init__()
        self.layers = nn.ModuleList([_EncoderLayer(d, dff, dropout) for _ in range(num_layers)])
        self.layer_norm = nn.LayerNorm(d)


class SelfAttentiveVAD(nn.Module):
    """Drop-in for vad.models.self_attention.SelfAttentiveVAD (eval / inference only).

    ``forward(features[B,T,F]) -> log-probs [B,T,2]``; extra keyword ``lengths`` applies the
    key-padding mask the reference encoder accepts (transformer.py:24-34, :432-447).
    ``compute_dtype``: "bf16" (default, tensor-core path) or "fp32" (<=1e-3 parity path).
    """

    def __init__(self, feature_size: int, num_layers: int, d_model: int, dropout: float,
                 compute_dtype: str = "bf16"):
        super().__init__()
        if d_model != 128:
            raise ValueError("the B200 kernels are specialised for d_model == 128")
        self.feature_size, self.num_layers, self.d_model = feature_size, num_layers, d_model
        self.compute_dtype = compute_dtype
        self.input_layer = nn.Sequential(nn.Linear(feature_size, d_model), nn.Identity(),
                                         nn.Dropout(dropout))
        self.encoder = _Encoder(num_layers, d_model, 4 * d_model, dropout)
        self.classifier = nn.Linear(d_model, 2)
        self.log_softmax = nn.LogSoftmax(dim=2)
        self._engine: Optional[VadEngine] = None
        self._engine_version = None
        self.eval()

    # weights are pushed to the library lazily and re-pushed when they change
    def _weights_version(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def engine(self, device: Optional[torch.device] = None) -> VadEngine:
        if device is None:
            p = next(self.parameters())
            device = p.device if p.is_cuda else torch.device("cuda")
        ver = (str(device), self._weights_version())
        if self._engine is None or self._engine.device != torch.device(
                "cuda", device.index if device.index is not None else torch.cuda.current_device()):
            if self._engine is not None:
                self._engine.close()
            self._engine = VadEngine(self.feature_size, self.num_layers, self.d_model,
                                     self.compute_dtype, device)
            self._engine_version = None
        if self._engine_version != ver:
            self._engine.load_state_dict(self.state_dict())
            self._engine_version = ver
        return self._engine

    def forward(self, features: Tensor, lengths: Optional[Tensor] = None) -> Tensor:
        if self.training:
            raise RuntimeError("vad_b200.SelfAttentiveVAD is inference-only (call .eval()); "
                               "training/backward is outside the accelerated path")
        eng = self.engine(features.device if features.is_cuda else None)
        _, logp = eng.forward(features, lengths, want_logp=True, want_prob=False)
        return logp

    def probabilities(self, features: Tensor, lengths: Optional[Tensor] = None) -> Tensor:
        """softmax(forward(x), -1)[..., 1] computed in the classifier kernel."""
        eng = self.engine(features.device if features.is_cuda else None)
        prob, _ = eng.forward(features, lengths, want_logp=False, want_prob=True)
        return prob


class ModelName(Enum):             # vad/models/model_factory.py:10-14
    DNN = "dnn"
    BDNN = "bdnn"
    ACAM = "acam"
    SELF_ATTENTIVE = "self-attention"


def create_model(model_config, feature_size: int, context_window_frames: int,
                 compute_dtype: str = "bf16"):
    """vad/models/model_factory.py:17-54; only the self-attention branch (:42-48) is on the
    accelerated path -- the DNN/bDNN/ACAM baselines are out of scope."""
    name = ModelName(model_config["name"])
    if name != ModelName.SELF_ATTENTIVE:
        raise NotImplementedError(
            f"model '{name.value}' is a comparison baseline of the reference and is outside the "
            "B200 hot path; only 'self-attention' is supported")
    sa = model_config["self_attention"]
    return SelfAttentiveVAD(feature_size, sa["num_layers"], sa["d_model"], sa["dropout"],
                            compute_dtype=compute_dtype)
