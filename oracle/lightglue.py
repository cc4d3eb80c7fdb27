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


from superslam_b200.lightglue_weights import make_random_weights, normalise_keys  # noqa: F401,E402


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


def self_block_debug(w, i, kpts, desc):
    """Every intermediate of SelfBlock i for one image (numpy), in the layouts the CUDA path stores:
    q (rotary applied, pre-scaled by 1/8) / k / v [H,N,64], logits S [H,N,N], ctx [N,256], msg [N,256],
    h1 = GELU(LayerNorm(fc1(cat[x,msg]))) [N,512], out x [N,256]."""
    w = {k: (v if isinstance(v, torch.Tensor) else torch.from_numpy(v)) for k, v in w.items()}
    x = torch.from_numpy(np.asarray(desc).astype(np.float32))
    enc = _posenc(w, torch.from_numpy(np.asarray(kpts, np.float32)))
    p = f"transformers.{i}.self_attn."
    with torch.no_grad():
        qkv = F.linear(x, w[p + "Wqkv.weight"], w[p + "Wqkv.bias"]).unflatten(-1, (N_HEADS, HEAD_DIM, 3)).transpose(0, 1)
        q, k, v = _rope(enc, qkv[..., 0]), _rope(enc, qkv[..., 1]), qkv[..., 2]
        S = torch.einsum("hid,hjd->hij", q, k) * HEAD_DIM ** -0.5
        ctx = torch.einsum("hij,hjd->hid", F.softmax(S, -1), v).transpose(0, 1).flatten(-2)
        msg = F.linear(ctx, w[p + "out_proj.weight"], w[p + "out_proj.bias"])
        cat = torch.cat([x, msg], -1)
        h1 = F.linear(cat, w[p + "ffn.0.weight"], w[p + "ffn.0.bias"])
        h1 = F.gelu(F.layer_norm(h1, (512,), w[p + "ffn.1.weight"], w[p + "ffn.1.bias"], 1e-5))
        out = x + F.linear(h1, w[p + "ffn.3.weight"], w[p + "ffn.3.bias"])
    return dict(q=(q * 0.125).numpy(), k=k.numpy(), v=v.numpy(), S=S.numpy(), ctx=ctx.numpy(), msg=msg.numpy(),
                h1=h1.numpy(), x=out.numpy(), cos=enc[0].numpy(), sin=enc[1].numpy())


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


# ---- fixed-precision restatement: the same model with every tensor the CUDA path keeps in fp16 rounded
# to fp16 where that path rounds it - GEMM operands (weights, the fp16 copy of the residual stream, q / k / v, the
# attention probabilities, the context, the FFN hidden activation) - and fp32 everywhere else (accumulation, softmax
# statistics, LayerNorm, the residual master, the assignment).  out_proj / to_out are folded into ffn.0 in fp64 and rounded
# once, as lightglue.cu does at weight load.  It is NOT bit-exact with the GPU (summation order inside the tensor cores,
# the polynomial exponentials), so index agreement with it is reported as a flip rate, not asserted as equality.
def _r16(t):
    return t.to(torch.float16).to(torch.float32)


def _ffn16(w, p, x16, ctx16, wo, bo):
    w1, b1 = w[p + "0.weight"].double(), w[p + "0.bias"].double()
    w1a, w1b = w1[:, :DIM], w1[:, DIM:]
    wf = torch.cat([w1a, w1b @ wo.double()], 1).float()
    bf = (b1 + w1b @ bo.double()).float()
    h = F.linear(torch.cat([x16, ctx16], -1), _r16(wf), bf)
    h = F.layer_norm(h, (h.shape[-1],), w[p + "1.weight"], w[p + "1.bias"], 1e-5)
    return F.linear(_r16(F.gelu(h)), _r16(w[p + "3.weight"]), w[p + "3.bias"])


def _attend16(q16, k16, v16, scale):
    s = torch.einsum("hid,hjd->hij", q16, k16) * scale
    e = torch.exp(s - s.max(-1, keepdim=True).values)
    ctx = torch.einsum("hij,hjd->hid", _r16(e), v16) / e.sum(-1, keepdim=True)
    return _r16(ctx.transpose(0, 1).flatten(-2))


def _self_block16(w, i, x, enc):
    p = f"transformers.{i}.self_attn."
    x16 = _r16(x)
    qkv = F.linear(x16, _r16(w[p + "Wqkv.weight"]), w[p + "Wqkv.bias"]).unflatten(-1, (N_HEADS, HEAD_DIM, 3)).transpose(0, 1)
    q, k, v = _r16(_rope(enc, qkv[..., 0])), _r16(_rope(enc, qkv[..., 1])), _r16(qkv[..., 2])
    ctx16 = _attend16(q, k, v, HEAD_DIM ** -0.5)
    return x + _ffn16(w, p + "ffn.", x16, ctx16, w[p + "out_proj.weight"], w[p + "out_proj.bias"])


def _cross_block16(w, i, x0, x1):
    p = f"transformers.{i}.cross_attn."
    heads = lambda t: t.unflatten(-1, (N_HEADS, HEAD_DIM)).transpose(0, 1)
    x016, x116 = _r16(x0), _r16(x1)
    wqk, wv = _r16(w[p + "to_qk.weight"]), _r16(w[p + "to_v.weight"])
    qk0, qk1 = _r16(heads(F.linear(x016, wqk, w[p + "to_qk.bias"]))), _r16(heads(F.linear(x116, wqk, w[p + "to_qk.bias"])))
    v0, v1 = _r16(heads(F.linear(x016, wv, w[p + "to_v.bias"]))), _r16(heads(F.linear(x116, wv, w[p + "to_v.bias"])))
    c0 = _attend16(qk0, qk1, v1, HEAD_DIM ** -0.5)
    c1 = _attend16(qk1, qk0, v0, HEAD_DIM ** -0.5)
    wo, bo = w[p + "to_out.weight"], w[p + "to_out.bias"]
    return x0 + _ffn16(w, p + "ffn.", x016, c0, wo, bo), x1 + _ffn16(w, p + "ffn.", x116, c1, wo, bo)


def log_assignment(w, i, x0, x1, return_sim=False):
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
    if return_sim:
        return s0 + s1 + cert, sim
    return s0 + s1 + cert


def filter_matches(scores, th=FILTER_THRESHOLD):
    max0, max1 = scores.max(1), scores.max(0)
    m0, m1 = max0.indices, max1.indices
    idx0 = torch.arange(m0.shape[0])
    mutual0 = idx0 == m1[m0]
    ms0 = torch.where(mutual0, max0.values.exp(), torch.zeros(()))
    valid0 = mutual0 & (ms0 > th)
    return torch.where(valid0, m0, torch.full_like(m0, -1)).to(torch.int32), ms0


def match(w, kpts0, desc0, kpts1, desc1, return_intermediates=False, fp16_storage=False):
    """kpts*: [N,2] f32 already normalised; desc*: [N,256] (fp16 values are widened to f32, as the
    engine's fp16 input binding would be).  Returns (matches0 int32 [N0], mscores0 f32 [N0]).
    fp16_storage: the fixed-precision restatement above (fp16 where the CUDA path stores fp16)."""
    w = {k: (v if isinstance(v, torch.Tensor) else torch.from_numpy(v)) for k, v in w.items()}
    k0 = torch.from_numpy(np.asarray(kpts0, np.float32))
    k1 = torch.from_numpy(np.asarray(kpts1, np.float32))
    x0 = torch.from_numpy(np.asarray(desc0).astype(np.float32))
    x1 = torch.from_numpy(np.asarray(desc1).astype(np.float32))
    inter = {}
    with torch.no_grad():
        e0, e1 = _posenc(w, k0), _posenc(w, k1)
        for i in range(N_LAYERS):
            x0 = _self_block16(w, i, x0, e0) if fp16_storage else _self_block(w, i, x0, e0)
            x1 = _self_block16(w, i, x1, e1) if fp16_storage else _self_block(w, i, x1, e1)
            if return_intermediates:
                inter[f"self{i}"] = (x0.numpy().copy(), x1.numpy().copy())
            x0, x1 = _cross_block16(w, i, x0, x1) if fp16_storage else _cross_block(w, i, x0, x1)
            if return_intermediates:
                inter[f"cross{i}"] = (x0.numpy().copy(), x1.numpy().copy())
        scores, sim = log_assignment(w, N_LAYERS - 1, x0, x1, return_sim=True)
        m0, ms0 = filter_matches(scores)
    if return_intermediates:
        inter["scores"] = scores.numpy()
        inter["sim"] = sim.numpy()
        return m0.numpy(), ms0.numpy(), inter
    return m0.numpy(), ms0.numpy()


def disagreement_report(scores: np.ndarray, m0: np.ndarray, ms0: np.ndarray, other_m0: np.ndarray,
                        other_ms0: np.ndarray | None = None, th: float = FILTER_THRESHOLD):
    """Classify every index where another implementation's matches0 (`other_m0`, e.g. the fp16 tensor-core path)
    differs from the oracle's (`m0`, from `scores` = the oracle's log-assignment matrix [N, M]).

    An index work path that is exact on ITS OWN scores can still disagree with the fp32 oracle where the oracle's decision
    hangs on a near-tie.  For each disagreeing query i (r = the oracle's row arg-max) the report gives the margins of the
    three decisions filter_matches takes (lightglue.py filter_matches; restated above):
      row_gap   S[i, r] - S[i, j']  when the other side matched another column j', else S[i, r] - second best of row i
      col_gap   | S[i, r] - best of column r over the other rows |        (mutual check m1[m0[i]] == i)
      thr_gap   | exp(S[i, r]) - th |                                      (mscores0 > th)
    and `margin` = the smallest margin that can explain the disagreement.  A disagreement is a legitimate flip when
    its margin is within the other path's score error; a large margin is a defect.
    `thr_gap_log` = | S[i, r] - log(th) | and `margin_log` is the same minimum with the threshold term in log-score
    units, so that it can be compared with an error measured on the log-assignment scores (explained_by_score_error).
    Returns dict(n, disagree, rows=[dict(i, oracle, other, kind, row_gap, col_gap, thr_gap, margin, thr_gap_log,
    margin_log)], max_margin, max_margin_log)."""
    S = np.asarray(scores, np.float64)
    n, m = S.shape
    rows = []
    for i in np.nonzero(np.asarray(m0) != np.asarray(other_m0))[0].tolist():
        r = int(np.argmax(S[i]))
        top = S[i, r]
        second = np.partition(S[i], -2)[-2] if m > 1 else -np.inf
        col = np.delete(S[:, r], i)
        col_gap = abs(top - col.max()) if col.size else np.inf
        thr_gap = abs(np.exp(top) - th)
        thr_gap_log = abs(top - np.log(th))
        j = int(other_m0[i])
        if j >= 0 and j != r:
            kind, row_gap = "row", top - S[i, j]
            margin = margin_log = row_gap          # the other side preferred j: only a row near-tie explains it
        else:
            kind, row_gap = ("validity", top - second)
            margin = min(row_gap, col_gap, thr_gap)
            margin_log = min(row_gap, col_gap, thr_gap_log)
        rows.append(dict(i=i, oracle=int(m0[i]), other=j, kind=kind, row_gap=float(row_gap), col_gap=float(col_gap),
                         thr_gap=float(thr_gap), margin=float(margin), thr_gap_log=float(thr_gap_log),
                         margin_log=float(margin_log)))
    return dict(n=n, disagree=len(rows), rows=rows, max_margin=max((r["margin"] for r in rows), default=0.0),
                max_margin_log=max((r["margin_log"] for r in rows), default=0.0))


def competitive_score_error(scores: np.ndarray, other_scores: np.ndarray, window: float = 3.0) -> float:
    """Largest |other - oracle| over the entries of the log-assignment matrix that can take part in a decision of
    filter_matches: those within `window` log units of their row's or their column's maximum (an entry far below both
    maxima decides nothing, and its absolute error grows with its distance from them)."""
    S = np.asarray(scores, np.float64)
    O = np.asarray(other_scores, np.float64)
    if S.size == 0:
        return 0.0
    near = (S >= S.max(1, keepdims=True) - window) | (S >= S.max(0, keepdims=True) - window)
    return float(np.abs(O - S)[near].max())


def explained_by_score_error(report: dict, score_error: float) -> list:
    """The rows of a disagreement_report that a score error of `score_error` (log units, both competitors of a decision
    may move by it) does NOT explain: empty for an implementation whose only differences are near-tie flips."""
    return [r for r in report["rows"] if not r["margin_log"] <= 2.0 * score_error]
