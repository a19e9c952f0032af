"""Seeded synthetic weights for benchmarks and smoke runs (the reference ships no checkpoints:
``.MISSING_LARGE_BLOBS`` lists every ``*.pth*``).  He-scaled normal conv weights, small biases, non-trivial
BatchNorm affine parameters and running statistics, in any of the three reference state-dict layouts."""
from __future__ import annotations

import torch

from .net import _GAUSS2, LAYER_NAMES, LAYER_SHAPES


def synthetic_state_dict(layout: str = "gauss2", seed: int = 0) -> dict:
    """layout in {'v1', 'bnflat', 'gauss2'} (SuperPointNet_pretrained / SuperPointNet / SuperPointNet_gauss2)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, (cin, cout, k) in zip(LAYER_NAMES, LAYER_SHAPES):
        cname, bname = _GAUSS2[name] if layout == "gauss2" else (name, "bn" + name[4:])
        sd[cname + ".weight"] = torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5
        sd[cname + ".bias"] = torch.randn(cout, generator=g) * 0.05
        if layout != "v1":
            sd[bname + ".weight"] = 1.0 + 0.1 * torch.randn(cout, generator=g)
            sd[bname + ".bias"] = 0.05 * torch.randn(cout, generator=g)
            sd[bname + ".running_mean"] = 0.05 * torch.randn(cout, generator=g)
            sd[bname + ".running_var"] = 1.0 + 0.2 * torch.rand(cout, generator=g)
            sd[bname + ".num_batches_tracked"] = torch.tensor(1)
    return sd
