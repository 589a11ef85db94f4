"""Synthetic weights and inputs for benchmarks, profiling tools and demos (product side: no dependency
on oracle/).  ``random_state`` fills the reference ``state_dict`` layout (vad/models/self_attention.py:12-21,
vad/modeling/transformer.py:10-61, :227-254, :366-375) with nn.Linear-style U(-1/sqrt(in), 1/sqrt(in))
values and slightly jittered LayerNorm affine terms; ``random_features`` draws log-mel-like inputs
(randn * 2 - 3).  Both are deterministic in ``seed`` and produce, bit for bit,
what the test oracle's own factories produce (tests/test_cabi_and_host.py checks it), so the CPU baseline
of bench.py and the GPU arm run the same model on the same data."""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict

import torch

from .engine import state_dict_keys


def state_shapes(feature_size: int, num_layers: int, d_model: int = 128) -> "OrderedDict[str, tuple]":
    d, dff = d_model, 4 * d_model            # d_ff = 4 * d_model: vad/models/self_attention.py:10
    shapes: "OrderedDict[str, tuple]" = OrderedDict()
    for k in state_dict_keys(num_layers):
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


def random_state(seed: int, feature_size: int = 64, num_layers: int = 3, d_model: int = 128,
                 ln_jitter: float = 0.1) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    shapes = state_shapes(feature_size, num_layers, d_model)
    st: Dict[str, torch.Tensor] = OrderedDict()
    for k, shp in shapes.items():
        if "layer_norm.weight" in k:
            st[k] = 1.0 + ln_jitter * (2 * torch.rand(shp, generator=g) - 1)
        elif "layer_norm.bias" in k:
            st[k] = ln_jitter * (2 * torch.rand(shp, generator=g) - 1)
        else:
            fan_in = shp[-1] if len(shp) == 2 else shapes[k.replace(".bias", ".weight")][-1]
            st[k] = (2 * torch.rand(shp, generator=g) - 1) * (1.0 / math.sqrt(fan_in))
        st[k] = st[k].to(torch.float32).contiguous()
    return st


def random_features(seed: int, B: int, T: int, feature_size: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B, T, feature_size, generator=g) * 2.0 - 3.0).to(torch.float32)
