"""LightGlue weight handling for the host side: key normalisation for cvg/LightGlue checkpoints,
conversion to the SSBW archive the C++ runtime loads, and seeded synthetic weights.

The reference ships no LightGlue weights (they are fetched by the un-vendored `lightglue` package,
/root/reference/utils/convert_lightglue_to_onnx.py:69); until a real checkpoint is dropped in
(tools/convert_lightglue_weights.py), tests and the benchmark run the correct architecture on the
synthetic weights below - throughput does not depend on the values.
"""
from __future__ import annotations

import math
import re
from collections import OrderedDict

import numpy as np
import torch

from .weights_io import save_archive

N_LAYERS = 9
DIM = 256
HEAD_DIM = 64


def normalise_keys(sd):
    """Accept both the in-memory names (transformers.{i}.self_attn.*) and the checkpoint-file names
    (self_attn.{i}.*, cross_attn.{i}.*), with or without a leading 'matcher.'."""
    out = OrderedDict()
    for k, v in sd.items():
        k = re.sub(r"^matcher\.", "", k)
        m = re.match(r"^(self_attn|cross_attn)\.(\d+)\.(.*)$", k)
        if m:
            k = f"transformers.{m.group(2)}.{m.group(1)}.{m.group(3)}"
        out[k] = v
    return out


def make_random_weights(seed: int = 7, sharpen: float = 2.0, matchability_bias: float = 3.0):
    """Seeded synthetic weights with torch.nn.Linear's default init (U(-1/sqrt(in), 1/sqrt(in))),
    LayerNorm weight ~1 / bias ~0, posenc.Wr ~ N(0,1) (gamma = 1.0).

    Two deliberate departures from a plain random init, so the assignment stage exercises matched,
    unmatched and thresholded branches instead of returning all -1: the last layer's final_proj is
    scaled by `sharpen` (peaked double softmax) and its matchability bias is raised by
    `matchability_bias` (logsigmoid ~ 0).  A real checkpoint needs neither.
    """
    g = torch.Generator().manual_seed(seed)

    def lin(out_f, in_f):
        b = 1.0 / math.sqrt(in_f)
        wt = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * b
        bs = (torch.rand(out_f, generator=g) * 2 - 1) * b
        return wt, bs

    sd = OrderedDict()
    sd["posenc.Wr.weight"] = torch.randn(HEAD_DIM // 2, 2, generator=g)
    for i in range(N_LAYERS):
        p = f"transformers.{i}.self_attn."
        sd[p + "Wqkv.weight"], sd[p + "Wqkv.bias"] = lin(3 * DIM, DIM)
        sd[p + "out_proj.weight"], sd[p + "out_proj.bias"] = lin(DIM, DIM)
        for blk in ("self_attn", "cross_attn"):
            q = f"transformers.{i}.{blk}.ffn."
            sd[q + "0.weight"], sd[q + "0.bias"] = lin(2 * DIM, 2 * DIM)
            sd[q + "1.weight"] = torch.ones(2 * DIM) + 0.1 * torch.randn(2 * DIM, generator=g)
            sd[q + "1.bias"] = 0.1 * torch.randn(2 * DIM, generator=g)
            sd[q + "3.weight"], sd[q + "3.bias"] = lin(DIM, 2 * DIM)
        p = f"transformers.{i}.cross_attn."
        sd[p + "to_qk.weight"], sd[p + "to_qk.bias"] = lin(DIM, DIM)
        sd[p + "to_v.weight"], sd[p + "to_v.bias"] = lin(DIM, DIM)
        sd[p + "to_out.weight"], sd[p + "to_out.bias"] = lin(DIM, DIM)
        p = f"log_assignment.{i}."
        sd[p + "matchability.weight"], sd[p + "matchability.bias"] = lin(1, DIM)
        sd[p + "final_proj.weight"], sd[p + "final_proj.bias"] = lin(DIM, DIM)
    last = f"log_assignment.{N_LAYERS - 1}."
    sd[last + "final_proj.weight"] = sd[last + "final_proj.weight"] * sharpen
    sd[last + "final_proj.bias"] = sd[last + "final_proj.bias"] * sharpen
    sd[last + "matchability.bias"] = sd[last + "matchability.bias"] + matchability_bias
    return OrderedDict((k, v.contiguous()) for k, v in sd.items())


def make_trained_like_weights(seed: int = 7, update_gain: float = 0.5, proj_gain: float = 2.2,
                              matchability_bias: float = 3.0, centre: bool = True):
    """Seeded synthetic weights whose assignment logits have the scale of a trained matcher instead of the ~50 of
    make_random_weights (whose sharpened final projection multiplies every upstream rounding error by that scale):
    the second FFN linear of every block is scaled by `update_gain` (residual updates about half the size of a default
    init, so the descriptors still dominate the residual stream) and the final projection is `proj_gain` x a random
    orthogonal matrix with zero bias, i.e. sim = proj_gain^2 / 16 * <x0, x1>: on the descriptor-like features of the
    tests the matched logits stand ~10 above the rest, the log-assignment scores of the competitive entries are O(1).
    Used where the north-star tolerances are asserted (mscores0 within 1e-3); a real checkpoint
    (tools/convert_lightglue_weights.py) replaces both synthetic sets."""
    sd = make_random_weights(seed, sharpen=1.0, matchability_bias=matchability_bias)
    g = torch.Generator().manual_seed(seed + 1)
    for i in range(N_LAYERS):
        for blk in ("self_attn", "cross_attn"):
            q = f"transformers.{i}.{blk}.ffn.3."
            sd[q + "weight"] = (sd[q + "weight"] * update_gain).contiguous()
            sd[q + "bias"] = (sd[q + "bias"] * update_gain).contiguous()
    qmat, _ = torch.linalg.qr(torch.randn(DIM, DIM, generator=g))
    last = f"log_assignment.{N_LAYERS - 1}."
    sd[last + "final_proj.weight"] = (qmat * proj_gain).contiguous()
    sd[last + "final_proj.bias"] = torch.zeros(DIM)
    if centre:
        # A random transformer adds the same bias-driven vector to every token, so <x0, x1> carries a large constant
        # (similarity logits ~50 of which ~10 discriminate).  A trained projection does not spend its dynamic range on
        # a constant: cancel the mean residual of a seeded calibration set in the projection's bias.
        mu = _calibration_mean(sd, seed)
        sd[last + "final_proj.bias"] = (-(sd[last + "final_proj.weight"] @ mu)).contiguous()
    return sd


def _calibration_mean(sd, seed: int) -> "torch.Tensor":
    """Mean residual vector after the 18 blocks on a seeded set of descriptor-like features (fp32, torch; the same
    arithmetic as oracle/lightglue.py, restated here because the product may not import the oracle)."""
    import torch.nn.functional as F

    g = torch.Generator().manual_seed(seed + 2)
    n = 192
    d0 = F.normalize(torch.randn(n, DIM, generator=g), dim=1)
    d1 = F.normalize(d0 + 0.05 * torch.randn(n, DIM, generator=g), dim=1)
    k0 = (torch.rand(n, 2, generator=g) - 0.5) * 1.5
    k1 = k0 - torch.tensor([[0.04, 0.0]])

    def rot(t):
        a, b = t.unflatten(-1, (-1, 2)).unbind(-1)
        return torch.stack((-b, a), -1).flatten(-2)

    def enc(k):
        p = (k @ sd["posenc.Wr.weight"].t()).repeat_interleave(2, -1)
        return p.cos(), p.sin()

    def heads(t):
        return t.unflatten(-1, (4, HEAD_DIM)).transpose(0, 1)

    def ffn(p, x, m):
        h = F.linear(torch.cat([x, m], -1), sd[p + "0.weight"], sd[p + "0.bias"])
        h = F.gelu(F.layer_norm(h, (2 * DIM,), sd[p + "1.weight"], sd[p + "1.bias"], 1e-5))
        return x + F.linear(h, sd[p + "3.weight"], sd[p + "3.bias"])

    with torch.no_grad():
        x, e = [d0, d1], [enc(k0), enc(k1)]
        for i in range(N_LAYERS):
            p = f"transformers.{i}.self_attn."
            for s_ in range(2):
                qkv = F.linear(x[s_], sd[p + "Wqkv.weight"], sd[p + "Wqkv.bias"]).unflatten(-1, (4, HEAD_DIM, 3)).transpose(0, 1)
                q, k, v = qkv[..., 0], qkv[..., 1], qkv[..., 2]
                c, s2 = e[s_]
                q, k = q * c + rot(q) * s2, k * c + rot(k) * s2
                ctx = F.softmax(q @ k.transpose(-1, -2) / 8.0, -1) @ v
                m = F.linear(ctx.transpose(0, 1).flatten(-2), sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])
                x[s_] = ffn(p + "ffn.", x[s_], m)
            p = f"transformers.{i}.cross_attn."
            qk = [heads(F.linear(t, sd[p + "to_qk.weight"], sd[p + "to_qk.bias"])) for t in x]
            vv = [heads(F.linear(t, sd[p + "to_v.weight"], sd[p + "to_v.bias"])) for t in x]
            sim = qk[0] @ qk[1].transpose(-1, -2) / 8.0
            m0 = F.softmax(sim, -1) @ vv[1]
            m1 = F.softmax(sim.transpose(-1, -2), -1) @ vv[0]
            ms = [F.linear(t.transpose(0, 1).flatten(-2), sd[p + "to_out.weight"], sd[p + "to_out.bias"]) for t in (m0, m1)]
            x = [ffn(p + "ffn.", x[s_], ms[s_]) for s_ in range(2)]
        return torch.cat(x).mean(0)


def save_state_dict(sd, path: str) -> None:
    """Write a (normalised) LightGlue state dict as an SSBW archive; token_confidence.* (unused with
    early exit disabled) and non-final log_assignment layers are dropped."""
    sd = normalise_keys(sd)
    out = OrderedDict()
    for k, v in sd.items():
        if k.startswith("token_confidence."):
            continue
        m = re.match(r"^log_assignment\.(\d+)\.", k)
        if m and int(m.group(1)) != N_LAYERS - 1:
            continue
        out[k] = (v.detach().cpu().float().numpy() if isinstance(v, torch.Tensor) else np.asarray(v, np.float32))
    save_archive(path, out)
