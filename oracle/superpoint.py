"""Oracle: SuperPoint dense network + host keypoint selection + descriptor gather (fp32, CPU).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows, line by line:
  * network          /root/reference/utils/convert_superpoint_to_onnx.py:38-64 (encoder), :72-90 (heads,
                     softmax, dustbin drop, depth-to-space, 9x9 NMS, descriptor L2 norm)
  * preprocess       /root/reference/src/SuperPoint.cc:768-778 (gray u8 -> f32 * (1/255))
  * select           /root/reference/src/SuperPoint.cc:696-719 (threshold, borders, sort, top-K, cells)
  * gather+normalise /root/reference/src/DescriptorGather.cu:14-56 (fp16 grid, fp32 tree sum, fp16 rows)
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

from superslam_b200.weights_io import load_archive

ENCODER = ["conv1a", "conv1b", "conv2a", "conv2b", "conv3a", "conv3b", "conv4a", "conv4b"]
POOL_AFTER = {"conv1b", "conv2b", "conv3b"}
NMS_RADIUS = 4  # baked at export time: convert_superpoint_to_onnx.py:97
DESC_DIM = 256


def load_weights(path: str) -> "OrderedDict[str, torch.Tensor]":
    return OrderedDict((k, torch.from_numpy(v)) for k, v in load_archive(path).items())


def _q16(x: torch.Tensor) -> torch.Tensor:
    """Round to fp16 and back (models fp16 storage with fp32 arithmetic)."""
    return x.half().float()


def preprocess(images_u8: np.ndarray) -> torch.Tensor:
    """[B,H,W] u8 -> [B,1,H,W] f32, value * float32(1/255) (cv::Mat::convertTo, SuperPoint.cc:776)."""
    x = torch.from_numpy(np.ascontiguousarray(images_u8)).float() * np.float32(1.0 / 255.0)
    return x[:, None]


def dense_forward(images_u8: np.ndarray, w, fp16_storage: bool = False):
    """Dense SuperPoint graph.  Returns (scores [B,8Hc,8Wc] f32 after NMS, descriptor grid
    [B,256,Hc,Wc] f32 L2-normalised over channels, raw softmax scores before NMS).

    fp16_storage=False is the reference graph in fp32 (the ground truth).  fp16_storage=True models
    the precision plan of the CUDA path: conv1a in fp32 on fp32 weights, every other layer on
    fp16-rounded weights, activations rounded to fp16 wherever they are stored (after every
    bias+ReLU(+pool)); accumulation, softmax, NMS and the channel norm stay fp32.
    """
    q = _q16 if fp16_storage else (lambda t: t)
    if isinstance(images_u8, np.ndarray) and images_u8.dtype == np.float32:   # an engine input [B,1,H,W], already / 255
        x = torch.from_numpy(np.ascontiguousarray(images_u8))
    else:
        x = preprocess(images_u8)
    with torch.no_grad():
        for name in ENCODER:
            wt = w[name + ".weight"] if name == "conv1a" else q(w[name + ".weight"])
            x = F.relu(F.conv2d(x, wt, w[name + ".bias"], padding=1))
            if name in POOL_AFTER:
                x = F.max_pool2d(x, 2, 2)
            x = q(x)
        pa = q(F.relu(F.conv2d(x, q(w["convPa.weight"]), w["convPa.bias"], padding=1)))
        logits = F.conv2d(pa, q(w["convPb.weight"]), w["convPb.bias"])
        prob = F.softmax(logits, 1)[:, :-1]
        b, _, hc, wc = prob.shape
        s = prob.permute(0, 2, 3, 1).reshape(b, hc, wc, 8, 8)
        s = s.permute(0, 1, 3, 2, 4).reshape(b, hc * 8, wc * 8)
        raw = s
        r = NMS_RADIUS
        s4 = s.unsqueeze(1)
        pooled = F.max_pool2d(s4, 2 * r + 1, stride=1, padding=r)
        scores = torch.where(s4 == pooled, s4, torch.zeros_like(s4)).squeeze(1)
        da = q(F.relu(F.conv2d(x, q(w["convDa.weight"]), w["convDa.bias"], padding=1)))
        d = F.conv2d(da, q(w["convDb.weight"]), w["convDb.bias"])
        d = F.normalize(d, p=2, dim=1)
    return scores.numpy(), d.numpy(), raw.numpy()


def dense_intermediates(images_u8: np.ndarray, w, fp16_storage: bool = True):
    """Per-layer activations of dense_forward (NHWC numpy, as the CUDA path stores them), for
    layer-by-layer parity checks.  Keys: conv1a..conv4b, convPa, convDa, logits, raw, grid."""
    q = _q16 if fp16_storage else (lambda t: t)
    x = preprocess(images_u8)
    out = {}
    with torch.no_grad():
        for name in ENCODER:
            wt = w[name + ".weight"] if name == "conv1a" else q(w[name + ".weight"])
            x = F.relu(F.conv2d(x, wt, w[name + ".bias"], padding=1))
            if name in POOL_AFTER:
                x = F.max_pool2d(x, 2, 2)
            x = q(x)
            out[name] = x.permute(0, 2, 3, 1).numpy().copy()
        pa = q(F.relu(F.conv2d(x, q(w["convPa.weight"]), w["convPa.bias"], padding=1)))
        da = q(F.relu(F.conv2d(x, q(w["convDa.weight"]), w["convDa.bias"], padding=1)))
        out["convPa"] = pa.permute(0, 2, 3, 1).numpy().copy()
        out["convDa"] = da.permute(0, 2, 3, 1).numpy().copy()
        logits = F.conv2d(pa, q(w["convPb.weight"]), w["convPb.bias"])
        out["logits"] = logits.permute(0, 2, 3, 1).numpy().copy()
        prob = F.softmax(logits, 1)[:, :-1]
        b, _, hc, wc = prob.shape
        s = prob.permute(0, 2, 3, 1).reshape(b, hc, wc, 8, 8).permute(0, 1, 3, 2, 4).reshape(b, hc * 8, wc * 8)
        out["raw"] = s.numpy().copy()
        d = F.normalize(F.conv2d(da, q(w["convDb.weight"]), w["convDb.bias"]), p=2, dim=1)
        out["grid"] = d.permute(0, 2, 3, 1).numpy().copy()
    return out


def nms_select(raw: np.ndarray, input_h: int, input_w: int, max_keypoints: int, keypoint_threshold: float,
               remove_borders: int):
    """9x9 NMS (convert_superpoint_to_onnx.py:82-87) on a given raw heat map [H',W'] followed by
    select_keypoints: the exact index work the CUDA nms/select kernels must reproduce bit for bit."""
    return select_keypoints(nms(raw), input_h, input_w, max_keypoints, keypoint_threshold, remove_borders,
                            raw.shape[0] // 8, raw.shape[1] // 8)


def nms(raw: np.ndarray) -> np.ndarray:
    """The in-graph 9x9 NMS alone (convert_superpoint_to_onnx.py:82-87): raw heat map [H',W'] -> score map."""
    s4 = torch.from_numpy(np.ascontiguousarray(raw))[None, None]
    pooled = F.max_pool2d(s4, 2 * NMS_RADIUS + 1, stride=1, padding=NMS_RADIUS)
    return torch.where(s4 == pooled, s4, torch.zeros_like(s4))[0, 0].numpy()


def select_keypoints(scores: np.ndarray, input_h: int, input_w: int, max_keypoints: int,
                     keypoint_threshold: float, remove_borders: int, desc_h: int, desc_w: int):
    """Host half of SuperPoint::select_and_gather (SuperPoint.cc:696-719) for one image.

    scores: [H', W'] f32.  Returns dict(xy [n,2] f32, score [n] f32, hw [n,2] i32, cell [n,2] i32).
    Order: score descending, ties by row descending then column descending (std::greater on
    pair<float,pair<int,int>>).  The threshold compare is float-promoted-to-double > double.
    """
    sh, sw = scores.shape
    rb = remove_borders
    hs, ws = np.nonzero(scores.astype(np.float64) > float(keypoint_threshold))
    keep = (hs >= rb) & (hs < sh - rb) & (ws >= rb) & (ws < sw - rb)
    hs, ws = hs[keep], ws[keep]
    sc = scores[hs, ws]
    # lexsort: last key is primary.  Descending on all three.
    order = np.lexsort((-ws, -hs, -sc.astype(np.float64)))
    order = order[:max_keypoints]
    hs, ws, sc = hs[order].astype(np.int32), ws[order].astype(np.int32), sc[order].astype(np.float32)
    scale_x = np.float32(input_w) / np.float32(sw)
    scale_y = np.float32(input_h) / np.float32(sh)
    xy = np.stack([ws.astype(np.float32) * scale_x, hs.astype(np.float32) * scale_y], 1).astype(np.float32)
    cell = np.stack([np.minimum(hs // 8, desc_h - 1), np.minimum(ws // 8, desc_w - 1)], 1).astype(np.int32)
    return dict(xy=xy.reshape(-1, 2), score=sc, hw=np.stack([hs, ws], 1).reshape(-1, 2), cell=cell.reshape(-1, 2))


def keypoint_disagreements(raw_ref: np.ndarray, raw_other: np.ndarray, hw_ref: np.ndarray, hw_other: np.ndarray,
                           max_keypoints: int, keypoint_threshold: float = 0.005, remove_borders: int = 4):
    """Index parity of another implementation's keypoint SET (`hw_other`, (row, col) on the score map, selected from
    its own heat map `raw_other`) against the oracle's (`hw_ref` from `raw_ref`, the reference graph's softmax heat
    map before NMS), decision by decision.  A pixel is a keypoint iff (SuperPoint.cc:696-719 after the graph's NMS,
    convert_superpoint_to_onnx.py:82-87)  score > threshold,  score >= every other pixel of its 9x9 window,  it is
    inside the border, and it ranks among the max_keypoints best.  For every pixel in the symmetric difference:
      selected by the oracle only:   margin = the smallest slack of those decisions in the oracle's map
      selected by the other only:    margin = the largest violation of them in the oracle's map
    and `bound` = 2 x the largest |raw_other - raw_ref| over the pixel's 9x9 window and the two candidates that sit
    at the top-K cut (the pixel's own score and its competitor's can each move by that much).  A disagreement with
    margin <= bound is a near-tie the other implementation's score error legitimately flips; anything else is a
    defect.  Returns dict(n_ref, n_other, differ, rows=[dict(hw, oracle_selected, score, margin, bound, why)],
    unexplained=[rows with margin > bound])."""
    R = NMS_RADIUS
    H, W = raw_ref.shape
    thr = float(keypoint_threshold)
    ref = set(map(tuple, np.asarray(hw_ref).reshape(-1, 2).tolist()))
    oth = set(map(tuple, np.asarray(hw_other).reshape(-1, 2).tolist()))
    cand = select_keypoints(nms(raw_ref), H, W, 1 << 30, keypoint_threshold, remove_borders, H // 8, W // 8)
    K = max_keypoints
    err = np.abs(np.asarray(raw_other, np.float64) - np.asarray(raw_ref, np.float64))
    s_last = s_next = None          # lowest selected / best excluded score of the oracle, when the top-K cut is active
    e_cut = 0.0
    if len(cand["score"]) > K:
        s_last, s_next = float(cand["score"][K - 1]), float(cand["score"][K])
        e_cut = float(max(err[tuple(cand["hw"][K - 1])], err[tuple(cand["hw"][K])]))
    pad = np.pad(np.asarray(raw_ref, np.float64), R, constant_values=-np.inf)
    pad_err = np.pad(err, R)
    rows = []
    for (h, w) in sorted(ref ^ oth):
        win = pad[h:h + 2 * R + 1, w:w + 2 * R + 1].copy()
        s = float(win[R, R])
        win[R, R] = -np.inf
        nb = float(win.max())
        e = float(pad_err[h:h + 2 * R + 1, w:w + 2 * R + 1].max())
        inside = remove_borders <= h < H - remove_borders and remove_borders <= w < W - remove_borders
        if (h, w) in ref:
            terms = {"threshold": s - thr, "nms": s - nb}
            if s_next is not None:
                terms["top-k"] = s - s_next
            why = min(terms, key=terms.get)
            margin = max(terms[why], 0.0)
        else:
            terms = {"threshold": thr - s, "nms": nb - s}
            if s_last is not None:
                terms["top-k"] = s_last - s
            why = max(terms, key=terms.get)
            margin = max(terms[why], 0.0) if inside else np.inf
        bound = 2.0 * max(e, e_cut if "top-k" in terms else 0.0)
        rows.append(dict(hw=(h, w), oracle_selected=(h, w) in ref, score=s, margin=float(margin), bound=bound, why=why))
    return dict(n_ref=len(ref), n_other=len(oth), differ=len(rows), rows=rows,
                unexplained=[r for r in rows if not r["margin"] <= r["bound"]])


def gather_normalize(grid_f16: np.ndarray, cell: np.ndarray) -> np.ndarray:
    """gather_normalize_kernel (DescriptorGather.cu:14-56).  grid_f16 [256,Hc,Wc] fp16, cell [n,2]
    (row, col).  fp32 sum of squares in the kernel's 256-wide tree order (strides 128..1),
    rsqrtf(sum + 1e-12f), output fp16 (round-to-nearest) rows [n,256]."""
    n = cell.shape[0]
    if n == 0:
        return np.zeros((0, grid_f16.shape[0]), np.float16)
    v = grid_f16[:, cell[:, 0], cell[:, 1]].T.astype(np.float32)  # [n,256]
    p = (v * v).astype(np.float32)
    stride = p.shape[1] // 2
    while stride > 0:
        p = (p[:, :stride] + p[:, stride:2 * stride]).astype(np.float32)
        stride //= 2
    inv = (np.float32(1.0) / np.sqrt(p[:, 0] + np.float32(1e-12), dtype=np.float32)).astype(np.float32)
    return (v * inv[:, None]).astype(np.float32).astype(np.float16)


def extract(images_u8: np.ndarray, w, max_keypoints: int, keypoint_threshold: float = 0.005,
            remove_borders: int = 4, fp16_storage: bool = False):
    """Full IFeatureExtractor::extract / extract_stereo restatement for a batch of same-size images.
    Returns a list (one per image) of dicts with xy, score, hw, cell, desc (fp16 [n,256]), plus the
    dense outputs for inspection."""
    b, h, wd = images_u8.shape
    scores, grid, raw = dense_forward(images_u8, w, fp16_storage)
    grid16 = grid.astype(np.float16)  # engine binding is fp16 (scripts/rebuild_engines.sh:92)
    out = []
    for i in range(b):
        k = select_keypoints(scores[i], h, wd, max_keypoints, keypoint_threshold, remove_borders,
                             grid.shape[2], grid.shape[3])
        k["desc"] = gather_normalize(grid16[i], k["cell"])
        k["scores_map"] = scores[i]
        k["grid_f16"] = grid16[i]
        k["raw_map"] = raw[i]
        out.append(k)
    return out
