"""Oracle: EigenPlaces global descriptor (ResNet18 + L2Norm/GeM/FC/L2Norm) + cosine retrieval (fp32, CPU).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows:
  * preprocess   /root/reference/src/EigenPlaces.cc:123-143  gray->RGB / BGR->RGB, cv::resize (bilinear),
                 convertTo(CV_32F, 1/255), per-channel ImageNet (v - mean) / std, HWC -> CHW
  * network      the un-vendored, un-pinned torch.hub model `gmberton/eigenplaces`
                 get_trained_model(backbone="ResNet18", fc_output_dim=512)
                 (/root/reference/utils/convert_eigenplaces_to_onnx.py:54-60): torchvision ResNet18 up to
                 layer4 (avgpool / fc dropped), then aggregation = L2Norm, GeM(p = 3 learnable, eps 1e-6),
                 Flatten, Linear(512 -> 512), L2Norm.  State-dict names as in that model: backbone.{0,1,4..7}.*,
                 aggregation.1.p, aggregation.3.{weight,bias}.
  * output       /root/reference/src/EigenPlaces.cc:145-174  fp32 row, cv::normalize(NORM_L2)
  * retrieval    /root/reference/src/PlaceRecognizer.cc:10-52 (CosineDescriptorIndex), :54-68
                 (TemporalConsistencyVoter); pinned by /root/reference/tests/test_place_recognizer.cc

Pinning: the resize port is checked against cv2.resize (the same OpenCV routine the reference calls), the
ResNet18 trunk against torchvision.models.resnet18 with shared weights, the index / voter against the
reference's own unit tests.  The trained weights are not available offline: PARITY UNPINNED for the network
values (architecture-correct, seeded synthetic weights), as for LightGlue.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32)   # EigenPlaces.cc:22
STD = np.array([0.229, 0.224, 0.225], dtype=np.float32)    # EigenPlaces.cc:23
BN_EPS = 1e-5
GEM_EPS = 1e-6
STAGES = [(4, 64, 1), (5, 128, 2), (6, 256, 2), (7, 512, 2)]   # backbone index, channels, first-block stride


# ---------------------------------------------------------------------------------------------------
# cv::resize(INTER_LINEAR) on 8-bit images (fixed point, 11 coefficient bits), restated
# ---------------------------------------------------------------------------------------------------
def _linear_coeffs(src: int, dst: int):
    """Per destination index: source index s0 (s1 = s0 + 1 clipped) and the two 11-bit weights."""
    scale = np.float64(src) / np.float64(dst)
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = f - s.astype(np.float32)
    return s, f


def resize_linear_u8(img: np.ndarray, dst_w: int, dst_h: int) -> np.ndarray:
    """cv::resize(src, dst, Size(dst_w, dst_h)) with the default INTER_LINEAR for CV_8UC1 / CV_8UC3."""
    src = img if img.ndim == 3 else img[:, :, None]
    h, w, cn = src.shape
    if (w, h) == (dst_w, dst_h):
        out = src.copy()
        return out if img.ndim == 3 else out[:, :, 0]
    if w == 2 * dst_w and h == 2 * dst_h:
        # exact 2x decimation is routed to the INTER_AREA fast path: rounded 2x2 mean
        s = src.astype(np.int32)
        out = (s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2
        out = out.astype(np.uint8)
        return out if img.ndim == 3 else out[:, :, 0]
    sx, fx = _linear_coeffs(w, dst_w)
    sy, fy = _linear_coeffs(h, dst_h)
    # horizontal: out-of-range taps are clamped with the fractional part forced to 0
    lo = sx < 0
    fx = np.where(lo, np.float32(0), fx)
    sx = np.where(lo, 0, sx)
    hi = sx >= w - 1
    fx = np.where(hi, np.float32(0), fx)
    sx = np.where(hi, w - 1, sx)
    a0 = _sat_short((np.float32(1) - fx) * np.float32(2048))
    a1 = _sat_short(fx * np.float32(2048))
    sx1 = np.minimum(sx + 1, w - 1)
    s32 = src.astype(np.int32)
    hbuf = s32[:, sx, :] * a0[None, :, None] + s32[:, sx1, :] * a1[None, :, None]   # [h, dst_w, cn]
    # vertical: rows are clipped, the weights are not
    b0 = _sat_short((np.float32(1) - fy) * np.float32(2048))
    b1 = _sat_short(fy * np.float32(2048))
    r0 = np.clip(sy, 0, h - 1)
    r1 = np.clip(sy + 1, 0, h - 1)
    S0 = hbuf[r0] >> 4
    S1 = hbuf[r1] >> 4
    out = (((b0[:, None, None] * S0) >> 16) + ((b1[:, None, None] * S1) >> 16) + 2) >> 2
    out = np.clip(out, 0, 255).astype(np.uint8)
    return out if img.ndim == 3 else out[:, :, 0]


def _sat_short(x: np.ndarray) -> np.ndarray:
    """cv::saturate_cast<short>(float): round half to even, saturate."""
    return np.clip(np.rint(x.astype(np.float64)), -32768, 32767).astype(np.int32)


def preprocess(image: np.ndarray, in_w: int, in_h: int) -> np.ndarray:
    """EigenPlaces::preprocess: u8 gray [H,W] or BGR [H,W,3] -> f32 [3, in_h, in_w] (RGB, normalised)."""
    if image.ndim == 2:
        rgb = np.repeat(image[:, :, None], 3, axis=2)          # COLOR_GRAY2RGB
    else:
        rgb = image[:, :, ::-1]                                 # COLOR_BGR2RGB
    rgb = resize_linear_u8(np.ascontiguousarray(rgb), in_w, in_h)
    v = rgb.astype(np.float32) * np.float32(1.0 / 255.0)        # convertTo(CV_32F, 1/255)
    v = (v - MEAN[None, None, :]) / STD[None, None, :]
    return np.ascontiguousarray(v.transpose(2, 0, 1))


# ---------------------------------------------------------------------------------------------------
# network
# ---------------------------------------------------------------------------------------------------
def _bn(x, w, prefix):
    return F.batch_norm(x, w[prefix + ".running_mean"], w[prefix + ".running_var"], w[prefix + ".weight"],
                        w[prefix + ".bias"], training=False, eps=BN_EPS)


def _basic_block(x, w, prefix, stride):
    """torchvision BasicBlock: conv3x3(s) bn relu conv3x3 bn (+ downsample(x)) add relu."""
    out = F.relu(_bn(F.conv2d(x, w[prefix + ".conv1.weight"], None, stride, 1), w, prefix + ".bn1"))
    out = _bn(F.conv2d(out, w[prefix + ".conv2.weight"], None, 1, 1), w, prefix + ".bn2")
    if (prefix + ".downsample.0.weight") in w:
        x = _bn(F.conv2d(x, w[prefix + ".downsample.0.weight"], None, stride, 0), w, prefix + ".downsample.1")
    return F.relu(out + x)


def backbone(x: torch.Tensor, w, taps=None) -> torch.Tensor:
    """[B,3,H,W] -> [B,512,H/32,W/32] (ResNet18 children()[:-2])."""
    x = F.relu(_bn(F.conv2d(x, w["backbone.0.weight"], None, 2, 3), w, "backbone.1"))
    if taps is not None:
        taps["stem"] = x
    x = F.max_pool2d(x, 3, 2, 1)
    if taps is not None:
        taps["pool"] = x
    for idx, _, stride in STAGES:
        for blk in range(2):
            x = _basic_block(x, w, f"backbone.{idx}.{blk}", stride if blk == 0 else 1)
        if taps is not None:
            taps[f"layer{idx - 3}"] = x
    return x


def aggregation(x: torch.Tensor, w) -> torch.Tensor:
    """L2Norm (channels) -> GeM -> Flatten -> Linear -> L2Norm."""
    x = F.normalize(x, p=2.0, dim=1)
    p = w["aggregation.1.p"]
    x = F.avg_pool2d(x.clamp(min=GEM_EPS).pow(p), (x.size(-2), x.size(-1))).pow(1.0 / p)
    x = x.flatten(1)
    x = F.linear(x, w["aggregation.3.weight"], w["aggregation.3.bias"])
    return F.normalize(x, p=2.0, dim=1)


def forward(w, x: np.ndarray, taps=None) -> np.ndarray:
    with torch.no_grad():
        return aggregation(backbone(torch.from_numpy(x), w, taps), w).numpy()


def compute_global_descriptor(w, image: np.ndarray, in_w: int, in_h: int) -> np.ndarray:
    """EigenPlaces::compute_global_descriptor: [1, 512] f32, L2-normalised (cv::normalize NORM_L2)."""
    d = forward(w, preprocess(image, in_w, in_h)[None])
    n = np.sqrt(np.sum(d.astype(np.float64) ** 2))
    return (d / np.float32(n) if n > 0 else d).astype(np.float32)


# ---------------------------------------------------------------------------------------------------
# retrieval (PlaceRecognizer.cc)
# ---------------------------------------------------------------------------------------------------
def normalized_row(desc: np.ndarray) -> np.ndarray:
    row = np.asarray(desc, dtype=np.float32).reshape(1, -1)
    n = float(np.sqrt(np.sum(row.astype(np.float64) ** 2)))   # cv::norm accumulates in double
    if n > 1e-12:
        row = (row / n).astype(np.float32)
    return row.copy()


class CosineDescriptorIndex:
    def __init__(self):
        self.ids = []
        self.db = np.zeros((0, 0), dtype=np.float32)

    def add(self, keyframe_id: int, desc: np.ndarray) -> None:
        row = normalized_row(desc)
        self.db = row if not self.ids else np.concatenate([self.db, row], axis=0)
        self.ids.append(int(keyframe_id))

    def size(self) -> int:
        return len(self.ids)

    def query(self, desc: np.ndarray, exclude_recent: int, top_k: int, min_score: float):
        """-> list of (keyframe_id, score), descending score (PlaceRecognizer.cc:26-52)."""
        M = len(self.ids)
        if M == 0 or M <= exclude_recent:
            return []
        q = normalized_row(desc)
        limit = M - exclude_recent
        scores = self.db[:limit] @ q[0]
        out = [(self.ids[i], float(scores[i])) for i in range(limit) if scores[i] >= np.float32(min_score)]
        out.sort(key=lambda c: -c[1])
        if top_k > 0 and len(out) > top_k:
            out = out[:top_k]
        return out


class TemporalConsistencyVoter:
    def __init__(self, required_votes: int, id_tolerance: int):
        self.required, self.tol = required_votes, id_tolerance
        self.streak, self.last_id, self.have_last = 0, 0, False

    def vote(self, best) -> bool:
        """best: (keyframe_id, score) or None (PlaceRecognizer.cc:54-68)."""
        if best is None:
            self.streak, self.have_last = 0, False
            return False
        kid = int(best[0])
        consistent = self.have_last and abs(kid - self.last_id) <= self.tol
        self.streak = self.streak + 1 if consistent else 1
        self.last_id, self.have_last = kid, True
        return self.streak >= self.required


def load_weights(path: str) -> "OrderedDict[str, torch.Tensor]":
    from superslam_b200.weights_io import load_archive

    return OrderedDict((k, torch.from_numpy(v)) for k, v in load_archive(path).items())
