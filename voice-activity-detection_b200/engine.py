"""VadEngine: one libvadb200 handle on one GPU; torch tensors in, torch tensors out.

PyTorch is plumbing here (device memory, streams); every FLOP of the forward pass runs in the
hand-written sm_100a kernels behind the C ABI.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Dict, Optional

import numpy as np
import torch

from . import _cabi


def state_dict_keys(num_layers: int):
    """Reference state_dict order (vad/models/self_attention.py:12-21; names in SURVEY.md 8 a1)."""
    keys = ["input_layer.0.weight", "input_layer.0.bias"]
    for l in range(num_layers):
        p = f"encoder.layers.{l}."
        for proj in ("query", "key", "value", "final"):
            keys += [p + f"self_attention.{proj}_projection.weight",
                     p + f"self_attention.{proj}_projection.bias"]
        keys += [p + "self_attention_sublayer.layer_norm.weight",
                 p + "self_attention_sublayer.layer_norm.bias",
                 p + "feed_forward.feed_forward.0.weight", p + "feed_forward.feed_forward.0.bias",
                 p + "feed_forward.feed_forward.3.weight", p + "feed_forward.feed_forward.3.bias",
                 p + "feed_forward_sublayer.layer_norm.weight",
                 p + "feed_forward_sublayer.layer_norm.bias"]
    keys += ["encoder.layer_norm.weight", "encoder.layer_norm.bias",
             "classifier.weight", "classifier.bias"]
    return keys


def infer_config(state: Dict[str, torch.Tensor]):
    d_model, feature_size = state["input_layer.0.weight"].shape
    L = 0
    while f"encoder.layers.{L}.self_attention.query_projection.weight" in state:
        L += 1
    return int(feature_size), L, int(d_model)


def pack_state(state: Dict[str, torch.Tensor], num_layers: int) -> torch.Tensor:
    """Flatten a reference state_dict into the packed fp32 blob vadb_load_weights takes."""
    parts = [state[k].detach().to(torch.float32).reshape(-1).cpu() for k in state_dict_keys(num_layers)]
    return torch.cat(parts).contiguous()


_DTYPES = {"fp32": _cabi.VADB_F32, "f32": _cabi.VADB_F32, "float32": _cabi.VADB_F32,
           "bf16": _cabi.VADB_BF16, "bfloat16": _cabi.VADB_BF16}


class HostTicket:
    """Handle of one asynchronous host call (VadEngine.forward_async)."""

    def __init__(self, engine, ticket, prob, logp, keep_alive):
        self._engine, self._ticket, self._prob, self._logp, self._keep = engine, ticket, prob, logp, keep_alive

    def wait(self):
        rc = self._engine._lib.vadb_host_wait(self._engine._h, self._ticket)
        _cabi.check(self._engine._lib, self._engine._h, rc, "vadb_host_wait")
        self._keep = None
        return self._prob, self._logp


class VadEngine:
    """Owns a vadb_handle.  ``compute_dtype``: "fp32" (<=1e-3 parity path) or "bf16"."""

    def __init__(self, feature_size: int, num_layers: int, d_model: int = 128,
                 compute_dtype: str = "bf16", device: Optional[torch.device] = None):
        self._lib = _cabi.load_library()
        if not torch.cuda.is_available():
            raise RuntimeError("vad_b200 needs a CUDA device (B200); there is no CPU fallback")
        device = torch.device(device if device is not None else "cuda")
        if device.type != "cuda":
            raise RuntimeError(f"vad_b200 runs on CUDA devices only, got {device}")
        self.device = torch.device("cuda", device.index if device.index is not None
                                   else torch.cuda.current_device())
        self.feature_size, self.num_layers, self.d_model = feature_size, num_layers, d_model
        self.compute_dtype = compute_dtype
        self._cfg = _cabi.VadbConfig(feature_size, num_layers, d_model, _DTYPES[compute_dtype])
        self._h = C.c_void_p()
        rc = self._lib.vadb_create(C.byref(self._h), C.byref(self._cfg), self.device.index)
        _cabi.check(self._lib, None, rc, "vadb_create")
        self.weight_count = int(self._lib.vadb_weight_count(C.byref(self._cfg)))
        self._async_out, self._async_n = {}, 0      # pinned output rings of forward_async, per shape

    # -- lifecycle ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.vadb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    @classmethod
    def from_state_dict(cls, state, compute_dtype="bf16", device=None):
        F_, L, d = infer_config(state)
        eng = cls(F_, L, d, compute_dtype, device)
        eng.load_state_dict(state)
        return eng

    def load_state_dict(self, state):
        self.load_blob(pack_state(state, self.num_layers))

    def load_blob(self, blob: torch.Tensor):
        """blob: packed fp32 weights, on the host or already on this device (e.g. the buffer a
        rank received from the one-off NCCL broadcast)."""
        assert blob.dtype == torch.float32 and blob.is_contiguous()
        on_dev = blob.is_cuda
        if on_dev:
            assert blob.device == self.device
        rc = self._lib.vadb_load_weights(self._h, C.c_void_p(blob.data_ptr()), blob.numel(),
                                         1 if on_dev else 0, self._stream_ptr())
        _cabi.check(self._lib, self._h, rc, "vadb_load_weights")

    def broadcast_weights(self, nccl_comm_ptr: int, root: int = 0):
        """C-level multi-GPU load: one ncclBroadcast of the packed blob from ``root``'s handle into this
        one (vadb_broadcast_weights).  ``nccl_comm_ptr`` is a raw ``ncclComm_t``.  Collective."""
        with torch.cuda.device(self.device):
            rc = self._lib.vadb_broadcast_weights(self._h, C.c_void_p(nccl_comm_ptr), root, self._stream_ptr())
        _cabi.check(self._lib, self._h, rc, "vadb_broadcast_weights")

    def _stream_ptr(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def reserve(self, B: int, T: int):
        _cabi.check(self._lib, self._h, self._lib.vadb_reserve(self._h, B, T), "vadb_reserve")

    @property
    def launch_count(self) -> int:
        return int(self._lib.vadb_launch_count(self._h))

    # -- hot path ----------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, lengths: Optional[torch.Tensor] = None,
                want_logp: bool = True, want_prob: bool = True):
        """x [B,T,F] fp32/bf16.  CUDA tensor -> async on the current stream, returns CUDA
        tensors; CPU tensor -> end-to-end host call (H2D, forward, D2H), returns CPU tensors.
        ``lengths`` [B]: key-padding mask j >= lengths[b].  Given on the HOST (CPU tensor, list, numpy) with CUDA
        features, the batch is run length-bucketed (``vadb_forward_ragged``: no work on the padding, outputs past
        a clip's processed length are 0); given as a CUDA tensor, as one padded batch like the reference.
        Returns (prob [B,T] fp32 or None, logp [B,T,2] fp32 or None)."""
        if x.dim() != 3 or x.shape[2] != self.feature_size:
            raise ValueError(f"expected [B,T,{self.feature_size}] features, got {tuple(x.shape)}")
        B, T, _ = x.shape
        if lengths is not None and isinstance(lengths, torch.Tensor):
            if lengths.numel() != B:
                raise ValueError("lengths must have one entry per clip")
        if not x.is_cuda:
            return self._forward_host(x, lengths, want_logp, want_prob)
        if x.device != self.device:
            raise ValueError(f"features on {x.device}, engine on {self.device}")
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.to(torch.float32)
        x = x.contiguous()
        ln_ptr = None
        prob = torch.empty((B, T), dtype=torch.float32, device=self.device) if want_prob else None
        logp = torch.empty((B, T, 2), dtype=torch.float32, device=self.device) if want_logp else None
        if lengths is not None and not (isinstance(lengths, torch.Tensor) and lengths.is_cuda):
            host_len = torch.as_tensor(lengths).to(device="cpu", dtype=torch.int32).contiguous()
            if host_len.numel() != B:
                raise ValueError("lengths must have one entry per clip")
            with torch.cuda.device(self.device):
                rc = self._lib.vadb_forward_ragged(
                    self._h, C.c_void_p(x.data_ptr()),
                    _cabi.VADB_BF16 if x.dtype == torch.bfloat16 else _cabi.VADB_F32,
                    C.c_void_p(host_len.data_ptr()), B, T,
                    C.c_void_p(prob.data_ptr()) if want_prob and prob.numel() else None,
                    C.c_void_p(logp.data_ptr()) if want_logp and logp.numel() else None,
                    self._stream_ptr())
            _cabi.check(self._lib, self._h, rc, "vadb_forward_ragged")
            return prob, logp
        if lengths is not None:
            lengths = lengths.to(device=self.device, dtype=torch.int32).contiguous()
            ln_ptr = C.c_void_p(lengths.data_ptr())
        with torch.cuda.device(self.device):
            rc = self._lib.vadb_forward(
                self._h, C.c_void_p(x.data_ptr()),
                _cabi.VADB_BF16 if x.dtype == torch.bfloat16 else _cabi.VADB_F32, ln_ptr, B, T,
                C.c_void_p(prob.data_ptr()) if want_prob and prob.numel() else None,
                C.c_void_p(logp.data_ptr()) if want_logp and logp.numel() else None,
                self._stream_ptr())
        _cabi.check(self._lib, self._h, rc, "vadb_forward")
        return prob, logp

    def _forward_host(self, x, lengths, want_logp, want_prob):
        B, T, _ = x.shape
        if x.dtype != torch.bfloat16:       # bf16 host features are uploaded as they are (half the bytes)
            x = x.to(torch.float32)
        x = x.contiguous()
        prob = torch.empty((B, T), dtype=torch.float32) if want_prob else None
        logp = torch.empty((B, T, 2), dtype=torch.float32) if want_logp else None
        ln_ptr = None
        if lengths is not None:
            lengths = torch.as_tensor(lengths).to(device="cpu", dtype=torch.int32).contiguous()
            ln_ptr = C.c_void_p(lengths.data_ptr())
        rc = self._lib.vadb_forward_host(
            self._h, C.c_void_p(x.data_ptr()),
            _cabi.VADB_BF16 if x.dtype == torch.bfloat16 else _cabi.VADB_F32, ln_ptr, B, T,
            C.c_void_p(prob.data_ptr()) if want_prob and prob.numel() else None,
            C.c_void_p(logp.data_ptr()) if want_logp and logp.numel() else None)
        _cabi.check(self._lib, self._h, rc, "vadb_forward_host")
        return prob, logp

    def forward_async(self, x: torch.Tensor, lengths: Optional[torch.Tensor] = None,
                      want_logp: bool = False, want_prob: bool = True) -> "HostTicket":
        """Streaming form of the host call: x [B,T,F] fp32 or bf16 in PINNED host memory; H2D, forward and
        D2H are only enqueued, so the upload of the next batch overlaps the compute of this one.  Returns a
        ticket whose ``wait()`` yields (prob, logp) pinned CPU tensors; the output buffers are a ring of
        four per shape: a result is valid until four further calls.  At most four calls may be
        outstanding (un-waited): a fifth raises RuntimeError without enqueueing anything."""
        if x.is_cuda or not x.is_pinned() or x.dtype not in (torch.float32, torch.bfloat16) or not x.is_contiguous():
            raise ValueError("forward_async needs a contiguous fp32/bf16 tensor in pinned host memory")
        if x.dim() != 3 or x.shape[2] != self.feature_size:
            raise ValueError(f"expected [B,T,{self.feature_size}] features, got {tuple(x.shape)}")
        B, T, _ = x.shape
        key = (B, T, want_logp, want_prob)
        ring = self._async_out.setdefault(key, [])
        slot = self._async_n % 4
        while len(ring) <= slot:
            ring.append((torch.empty((B, T), dtype=torch.float32).pin_memory() if want_prob else None,
                         torch.empty((B, T, 2), dtype=torch.float32).pin_memory() if want_logp else None))
        prob, logp = ring[slot]
        ln_ptr, keep = None, None
        if lengths is not None:
            keep = torch.as_tensor(lengths).to(device="cpu", dtype=torch.int32).contiguous()
            ln_ptr = C.c_void_p(keep.data_ptr())
        ticket = C.c_long(-1)
        rc = self._lib.vadb_forward_host_async(
            self._h, C.c_void_p(x.data_ptr()),
            _cabi.VADB_BF16 if x.dtype == torch.bfloat16 else _cabi.VADB_F32, ln_ptr, B, T,
            C.c_void_p(prob.data_ptr()) if want_prob and prob.numel() else None,
            C.c_void_p(logp.data_ptr()) if want_logp and logp.numel() else None, C.byref(ticket))
        _cabi.check(self._lib, self._h, rc, "vadb_forward_host_async")
        self._async_n += 1
        return HostTicket(self, ticket.value, prob, logp, (x, keep))

    def predict_probabilities(self, feature, half: int, jump: int):
        """feature [L,F] (numpy / CPU tensor -> host call; CUDA tensor -> device call).
        Returns (probs [L,W], mean [L]) as numpy arrays (host) or CUDA tensors (device)."""
        W = 2 * (half - 1) // jump + 3
        if isinstance(feature, torch.Tensor) and feature.is_cuda:
            if feature.dim() != 2 or feature.shape[1] != self.feature_size:
                raise ValueError(f"expected [L,{self.feature_size}] features, got {tuple(feature.shape)}")
            if feature.device != self.device:
                raise ValueError(f"features on {feature.device}, engine on {self.device}")
            feat = feature.to(torch.float32).contiguous()
            L = feat.shape[0]
            probs = torch.empty((L, W), dtype=torch.float32, device=self.device)
            mean = torch.empty((L,), dtype=torch.float32, device=self.device)
            with torch.cuda.device(self.device):
                rc = self._lib.vadb_predict_probabilities(
                    self._h, C.c_void_p(feat.data_ptr()), L, half, jump,
                    C.c_void_p(probs.data_ptr()) if L else None,
                    C.c_void_p(mean.data_ptr()) if L else None, self._stream_ptr())
            _cabi.check(self._lib, self._h, rc, "vadb_predict_probabilities")
            return probs, mean
        feat = np.ascontiguousarray(np.asarray(feature, dtype=np.float32))
        if feat.ndim != 2 or feat.shape[1] != self.feature_size:
            raise ValueError(f"expected [L,{self.feature_size}] features, got {feat.shape}")
        L = feat.shape[0]
        probs = np.empty((L, W), dtype=np.float32)
        mean = np.empty((L,), dtype=np.float32)
        rc = self._lib.vadb_predict_probabilities_host(
            self._h, feat.ctypes.data_as(C.c_void_p), L, half, jump,
            probs.ctypes.data_as(C.c_void_p), mean.ctypes.data_as(C.c_void_p))
        _cabi.check(self._lib, self._h, rc, "vadb_predict_probabilities_host")
        return probs, mean

    def logmel(self, audio, sample_rate: int, n_fft: int, hop: int, win: int, n_mels: int):
        """Log-mel frames [L, n_mels] of mono PCM (vad/acoustics/transforms/log_mel_spectrogram.py:19-32
        + feature_extractor.py:77-80) computed on the device.  ``audio``: 1-D CUDA tensor -> CUDA
        tensor; numpy / CPU tensor -> uploaded, result returned as numpy."""
        on_dev = isinstance(audio, torch.Tensor) and audio.is_cuda
        a = audio if on_dev else torch.as_tensor(np.ascontiguousarray(audio, dtype=np.float32)).to(self.device)
        a = a.to(torch.float32).contiguous()
        n = a.numel()
        L = int(self._lib.vadb_logmel_frames(n, hop))
        feat = torch.empty((L, n_mels), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            rc = self._lib.vadb_logmel(self._h, C.c_void_p(a.data_ptr()), n, sample_rate, n_fft, hop, win,
                                       n_mels, C.c_void_p(feat.data_ptr()), self._stream_ptr())
        _cabi.check(self._lib, self._h, rc, "vadb_logmel")
        return feat if on_dev else feat.cpu().numpy()

    def predict_audio(self, audio, sample_rate: int, n_fft: int, hop: int, win: int, half: int, jump: int,
                      want_features: bool = False):
        """Host PCM -> (probs [L,W], mean [L], features [L,F] or None), everything between on the
        device in one call (vad/predictor.py:159-262 including the feature extraction of :160)."""
        a = np.ascontiguousarray(np.asarray(audio, dtype=np.float32))
        if a.ndim != 1:
            raise ValueError("expected mono PCM [n_samples]")
        W = 2 * (half - 1) // jump + 3
        L = int(self._lib.vadb_logmel_frames(a.shape[0], hop))
        probs = np.empty((L, W), dtype=np.float32)
        mean = np.empty((L,), dtype=np.float32)
        feat = np.empty((L, self.feature_size), dtype=np.float32) if want_features else None
        rc = self._lib.vadb_predict_audio_host(
            self._h, a.ctypes.data_as(C.c_void_p), a.shape[0], sample_rate, n_fft, hop, win, half, jump,
            feat.ctypes.data_as(C.c_void_p) if want_features else None,
            probs.ctypes.data_as(C.c_void_p), mean.ctypes.data_as(C.c_void_p))
        _cabi.check(self._lib, self._h, rc, "vadb_predict_audio_host")
        return probs, mean, feat

    def attention(self, q, k, v, lengths=None):
        """Stage-level entry (kernel parity tests / roofline bench): q,k,v [B,T,128] CUDA,
        fp32 -> CUDA-core kernel, bf16 -> tcgen05 kernel."""
        assert q.is_cuda and q.shape == k.shape == v.shape and q.shape[-1] == 128
        assert q.dtype == k.dtype == v.dtype and q.dtype in (torch.float32, torch.bfloat16)
        if not (q.device == k.device == v.device == self.device):
            raise ValueError(f"q/k/v must be on the engine's device {self.device}")
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        B, T, _ = q.shape
        o = torch.empty_like(q)
        ln_ptr = None
        if lengths is not None:
            lengths = lengths.to(device=self.device, dtype=torch.int32).contiguous()
            ln_ptr = C.c_void_p(lengths.data_ptr())
        with torch.cuda.device(self.device):
            rc = self._lib.vadb_attention(
                self._h, C.c_void_p(q.data_ptr()), C.c_void_p(k.data_ptr()),
                C.c_void_p(v.data_ptr()), C.c_void_p(o.data_ptr()),
                _cabi.VADB_BF16 if q.dtype == torch.bfloat16 else _cabi.VADB_F32, ln_ptr, B, T,
                self._stream_ptr())
        _cabi.check(self._lib, self._h, rc, "vadb_attention")
        return o

    def positional_table(self, T: int) -> np.ndarray:
        out = np.empty((T, 128), dtype=np.float32)
        rc = self._lib.vadb_positional_table(self._h, T, out.ctypes.data_as(C.c_void_p))
        _cabi.check(self._lib, self._h, rc, "vadb_positional_table")
        return out
