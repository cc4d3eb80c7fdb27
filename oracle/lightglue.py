"""Oracle: LightGlue matcher (features="superpoint"), fp32 torch on CPU.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: the reference delegates the
arithmetic to the un-vendored, un-pinned package cvg/LightGlue
(/root/reference/utils/convert_lightglue_to_onnx.py:4-9,53-54,69) and ships no weights.  This file
restates that package's published model (lightglue/lightglue.py, v0.0 "superpoint_lightglue")
under the export wrapper's settings (convert_lightglue_to_onnx.py:61,69-75,88-89):

  * keypoint normalisation is done by the caller (src/LightGlue.cc:241-251); the graph's own is a no-op
  * n_layers 9, heads 4, dim 256, input_proj = identity, flash off
  * depth_confidence = width_confidence = -1  -> no early exit, no pruning; token_confidence unused
  * filter_threshold 0.1; outputs matches0 (int32, -1 = unmatched) and matching_scores0

State-dict key names are cvg's:  posenc.Wr.weight;  transformers.{i}.self_attn.{Wqkv,out_proj}.*,
transformers.{i}.self_attn.ffn.{0,1,3}.*;  transformers.{i}.cross_attn.{to_qk,to_v,to_out}.*,
transformers.{i}.cross_attn.ffn.{0,1,3}.*;  log_assignment.{i}.{matchability,final_proj}.*.
(Upstream checkpoint files name the blocks self_attn.{i}.* / cross_attn.{i}.* and are renamed on
load; normalise_keys() accepts both.)
"""
from __future__ import annotations

import math
import re
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

N_LAYERS = 9
N_HEADS = 4
DIM = 256
HEAD_DIM = 64
FILTER_THRESHOLD = 0.1


def normalise_keys(sd):
    out = OrderedDict()
    for k, v in sd.items():
        k = re.sub(r"^matcher\.", "", k)
        m = re.match(r"^(self_attn|cross_attn)\.(\d+)\.(.*)$", k)
        if m:
            k = f"transformers.{m.group(2)}.{m.group(1)}.{m.group(3)}"
        out[k] = v
    return out


def make_random_weights(seed: int = 7, sharpen: float = 6.0, matchability_bias: float = 3.0):
    """Seeded synthetic weights with torch.nn.Linear's default init (U(-1/sqrt(in), 1/sqrt(in))),
    LayerNorm weight 1 / bias 0, posenc.Wr ~ N(0,1) (gamma = 1.0).

    Two deliberate departures from a plain random init, so the assignment stage exercises matched,
    unmatched and thresholded branches instead of returning all -1: the last layer's final_proj is
    scaled by `sharpen` (peaked double softmax) and its matchability bias is raised to
    `matchability_bias` (logsigmoid ~ 0).  The real checkpoint needs neither.
    """
    g = torch.Generator().manual_seed(seed)

    def lin(out_f, in_f, bias=True):
        b = 1.0 / math.sqrt(in_f)
        wt = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * b
        bs = (torch.rand(out_f, generator=g) * 2 - 1) * b if bias else None
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
    # reorder keys so that weight/bias pairs are adjacent and deterministic
    return OrderedDict((k, v.contiguous()) for k, v in sd.items())


def normalize_keypoints(xy: np.ndarray, image_width: int, image_height: int) -> np.ndarray:
    """LightGlue::store_keypoints (src/LightGlue.cc:241-251), fp32 host arithmetic."""
    scale = np.float32(max(image_width, image_height)) / np.float32(2.0)
    cx = np.float32(image_width) / np.float32(2.0)
    cy = np.float32(image_height) / np.float32(2.0)
    out = np.empty_like(xy, dtype=np.float32)
    out[:, 0] = (xy[:, 0].astype(np.float32) - cx) / scale
    out[:, 1] = (xy[:, 1].astype(np.float32) - cy) / scale
    return out


def _posenc(w, kpts):  # kpts [N,2] normalised -> (cos, sin) each [N,64]
    proj = kpts @ w["posenc.Wr.weight"].t()  # [N,32]
    cos, sin = torch.cos(proj), torch.sin(proj)
    return cos.repeat_interleave(2, -1), sin.repeat_interleave(2, -1)


def _rotate_half(x):
    x = x.unflatten(-1, (-1, 2))
    x1, x2 = x.unbind(-1)
    return torch.stack((-x2, x1), -1).flatten(-2)


def _rope(enc, t):  # t [H,N,64]
    return t * enc[0] + _rotate_half(t) * enc[1]


def _ffn(w, p, x):
    h = F.linear(x, w[p + "0.weight"], w[p + "0.bias"])
    h = F.layer_norm(h, (h.shape[-1],), w[p + "1.weight"], w[p + "1.bias"], 1e-5)
    h = F.gelu(h)
    return F.linear(h, w[p + "3.weight"], w[p + "3.bias"])


def _self_block(w, i, x, enc):
    p = f"transformers.{i}.self_attn."
    qkv = F.linear(x, w[p + "Wqkv.weight"], w[p + "Wqkv.bias"])  # [N,768]
    qkv = qkv.unflatten(-1, (N_HEADS, HEAD_DIM, 3)).transpose(0, 1)  # [H,N,64,3]
    q, k, v = qkv[..., 0], qkv[..., 1], qkv[..., 2]
    q, k = _rope(enc, q), _rope(enc, k)
    s = HEAD_DIM ** -0.5
    attn = F.softmax(torch.einsum("hid,hjd->hij", q, k) * s, -1)
    ctx = torch.einsum("hij,hjd->hid", attn, v)
    msg = F.linear(ctx.transpose(0, 1).flatten(-2), w[p + "out_proj.weight"], w[p + "out_proj.bias"])
    return x + _ffn(w, p + "ffn.", torch.cat([x, msg], -1))


def _cross_block(w, i, x0, x1):
    p = f"transformers.{i}.cross_attn."
    heads = lambda t: t.unflatten(-1, (N_HEADS, HEAD_DIM)).transpose(0, 1)
    qk0 = heads(F.linear(x0, w[p + "to_qk.weight"], w[p + "to_qk.bias"]))
    qk1 = heads(F.linear(x1, w[p + "to_qk.weight"], w[p + "to_qk.bias"]))
    v0 = heads(F.linear(x0, w[p + "to_v.weight"], w[p + "to_v.bias"]))
    v1 = heads(F.linear(x1, w[p + "to_v.weight"], w[p + "to_v.bias"]))
    sc = (HEAD_DIM ** -0.5) ** 0.5
    sim = torch.einsum("hid,hjd->hij", qk0 * sc, qk1 * sc)
    attn01 = F.softmax(sim, -1)
    attn10 = F.softmax(sim.transpose(-2, -1).contiguous(), -1)
    m0 = torch.einsum("hij,hjd->hid", attn01, v1)
    m1 = torch.einsum("hji,hjd->hid", attn10.transpose(-2, -1), v0)
    m0 = F.linear(m0.transpose(0, 1).flatten(-2), w[p + "to_out.weight"], w[p + "to_out.bias"])
    m1 = F.linear(m1.transpose(0, 1).flatten(-2), w[p + "to_out.weight"], w[p + "to_out.bias"])
    x0 = x0 + _ffn(w, p + "ffn.", torch.cat([x0, m0], -1))
    x1 = x1 + _ffn(w, p + "ffn.", torch.cat([x1, m1], -1))
    return x0, x1


def log_assignment(w, i, x0, x1):
    """MatchAssignment + sigmoid_log_double_softmax, inner [N,M] block only (the dustbin row/column
    never enters filter_matches)."""
    p = f"log_assignment.{i}."
    md0 = F.linear(x0, w[p + "final_proj.weight"], w[p + "final_proj.bias"]) / DIM ** 0.25
    md1 = F.linear(x1, w[p + "final_proj.weight"], w[p + "final_proj.bias"]) / DIM ** 0.25
    sim = md0 @ md1.t()
    z0 = F.linear(x0, w[p + "matchability.weight"], w[p + "matchability.bias"])  # [N,1]
    z1 = F.linear(x1, w[p + "matchability.weight"], w[p + "matchability.bias"])
    cert = F.logsigmoid(z0) + F.logsigmoid(z1).t()
    s0 = F.log_softmax(sim, 1)
    s1 = F.log_softmax(sim.t().contiguous(), 1).t()
    return s0 + s1 + cert


def filter_matches(scores, th=FILTER_THRESHOLD):
    max0, max1 = scores.max(1), scores.max(0)
    m0, m1 = max0.indices, max1.indices
    idx0 = torch.arange(m0.shape[0])
    mutual0 = idx0 == m1[m0]
    ms0 = torch.where(mutual0, max0.values.exp(), torch.zeros(()))
    valid0 = mutual0 & (ms0 > th)
    return torch.where(valid0, m0, torch.full_like(m0, -1)).to(torch.int32), ms0


def match(w, kpts0, desc0, kpts1, desc1, return_intermediates=False):
    """kpts*: [N,2] f32 already normalised; desc*: [N,256] (fp16 values are widened to f32, as the
    engine's fp16 input binding would be).  Returns (matches0 int32 [N0], mscores0 f32 [N0])."""
    w = {k: (v if isinstance(v, torch.Tensor) else torch.from_numpy(v)) for k, v in w.items()}
    k0 = torch.from_numpy(np.asarray(kpts0, np.float32))
    k1 = torch.from_numpy(np.asarray(kpts1, np.float32))
    x0 = torch.from_numpy(np.asarray(desc0).astype(np.float32))
    x1 = torch.from_numpy(np.asarray(desc1).astype(np.float32))
    inter = {}
    with torch.no_grad():
        e0, e1 = _posenc(w, k0), _posenc(w, k1)
        for i in range(N_LAYERS):
            x0 = _self_block(w, i, x0, e0)
            x1 = _self_block(w, i, x1, e1)
            if return_intermediates:
                inter[f"self{i}"] = (x0.numpy().copy(), x1.numpy().copy())
            x0, x1 = _cross_block(w, i, x0, x1)
            if return_intermediates:
                inter[f"cross{i}"] = (x0.numpy().copy(), x1.numpy().copy())
        scores = log_assignment(w, N_LAYERS - 1, x0, x1)
        m0, ms0 = filter_matches(scores)
    if return_intermediates:
        inter["scores"] = scores.numpy()
        return m0.numpy(), ms0.numpy(), inter
    return m0.numpy(), ms0.numpy()
