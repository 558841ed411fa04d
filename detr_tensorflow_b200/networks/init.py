"""Seeded synthetic initialisation in the reference's layouts (the pretrained checkpoint of networks/weights.py:5-11
is an online download and a TF object-graph file: out of scope).

He-style conv init with FrozenBN weight ~ 1, bias ~ 0, running_var ~ 1 keeps activations O(1) through the 50-layer
backbone; Linear layers use the reference's Glorot-uniform (custom_layers.py:43-44)."""
import math
from collections import OrderedDict

import torch

from .spec import model_params


def init_params(seed=0, **kw):
    g = torch.Generator().manual_seed(seed)
    out = OrderedDict()
    for name, p in model_params(**kw).items():
        shape, kind = p.shape, p.kind
        if kind == "conv":
            kh, kw_, ci, co = shape
            fan_in = kh * kw_ * ci
            std = math.sqrt(2.0 / fan_in)
            if name.endswith("conv3/kernel"):
                std *= 0.5
            if name.startswith("input_proj"):
                std = math.sqrt(1.0 / fan_in)
            t = torch.randn(shape, generator=g, dtype=torch.float64) * std
        elif kind == "bn_w":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g, dtype=torch.float64)
        elif kind in ("bn_b", "bn_mean"):
            t = 0.05 * torch.randn(shape, generator=g, dtype=torch.float64)
        elif kind == "bn_var":
            t = 1.0 + 0.1 * torch.rand(shape, generator=g, dtype=torch.float64)
        elif kind in ("linear_w", "embed", "dense_w"):
            lim = math.sqrt(6.0 / (shape[0] + shape[1]))
            t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * lim
        elif kind == "linear_b":
            t = 0.02 * torch.randn(shape, generator=g, dtype=torch.float64)
        elif kind == "ln_g":
            t = 1.0 + 0.05 * torch.randn(shape, generator=g, dtype=torch.float64)
        elif kind == "ln_b":
            t = 0.02 * torch.randn(shape, generator=g, dtype=torch.float64)
        else:
            raise ValueError(kind)
        out[name] = t.to(torch.float32)
    return out
