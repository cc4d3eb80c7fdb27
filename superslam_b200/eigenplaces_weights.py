"""EigenPlaces (ResNet18, 512-d) weight handling for the host side: seeded synthetic weights with the
state-dict names of the reference's torch.hub model and conversion to the SSBW archive the C++ runtime loads.

The reference fetches the trained model through torch.hub at export time
(/root/reference/utils/convert_eigenplaces_to_onnx.py:54-60) and ships no weights; until a real
state dict is dropped in (save_state_dict accepts it as is), tests and the benchmark run the correct
architecture on the synthetic weights below.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch

from .weights_io import save_archive

STAGES = [(4, 64, 64, 1), (5, 64, 128, 2), (6, 128, 256, 2), (7, 256, 512, 2)]   # index, cin, cout, stride
DESC_DIM = 512


def make_random_weights(seed: int = 11) -> "OrderedDict[str, torch.Tensor]":
    """torchvision's ResNet init (kaiming-normal fan_out convolutions) with non-trivial BatchNorm
    statistics, so that folding BN into the convolutions is exercised; the second BN of every block is
    damped (gamma ~ 0.5) to keep the residual stream O(1) through the eight blocks."""
    g = torch.Generator().manual_seed(seed)

    def conv(cout, cin, k):
        std = math.sqrt(2.0 / (cout * k * k))
        return torch.randn(cout, cin, k, k, generator=g) * std

    def bn(sd, prefix, c, gamma_mid=1.0):
        sd[prefix + ".weight"] = gamma_mid * (0.6 + 0.8 * torch.rand(c, generator=g))
        sd[prefix + ".bias"] = 0.1 * torch.randn(c, generator=g)
        sd[prefix + ".running_mean"] = 0.1 * torch.randn(c, generator=g)
        sd[prefix + ".running_var"] = 0.5 + torch.rand(c, generator=g)

    sd = OrderedDict()
    sd["backbone.0.weight"] = conv(64, 3, 7)
    bn(sd, "backbone.1", 64)
    for idx, cin, cout, stride in STAGES:
        for blk in range(2):
            p = f"backbone.{idx}.{blk}"
            c_in = cin if blk == 0 else cout
            sd[p + ".conv1.weight"] = conv(cout, c_in, 3)
            bn(sd, p + ".bn1", cout)
            sd[p + ".conv2.weight"] = conv(cout, cout, 3)
            bn(sd, p + ".bn2", cout, 0.5)
            if blk == 0 and (stride != 1 or c_in != cout):
                sd[p + ".downsample.0.weight"] = conv(cout, c_in, 1)
                bn(sd, p + ".downsample.1", cout)
    sd["aggregation.1.p"] = torch.tensor([3.0])
    b = 1.0 / math.sqrt(512)
    sd["aggregation.3.weight"] = (torch.rand(DESC_DIM, 512, generator=g) * 2 - 1) * b
    sd["aggregation.3.bias"] = (torch.rand(DESC_DIM, generator=g) * 2 - 1) * b
    return sd


def save_state_dict(sd, path: str) -> None:
    """Write a (real or synthetic) EigenPlaces state dict as an SSBW archive; `num_batches_tracked`
    counters are dropped."""
    out = OrderedDict()
    for k, v in sd.items():
        if k.endswith("num_batches_tracked"):
            continue
        a = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
        out[k] = np.ascontiguousarray(a, dtype=np.float32).reshape(a.shape if a.ndim else (1,))
    save_archive(path, out)
