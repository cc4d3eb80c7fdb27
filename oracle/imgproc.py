"""Oracle: the image front door either side of the networks, numpy.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
  * remap_linear_u8     cv::remap(img, out, M1, M2, cv::INTER_LINEAR) with CV_32FC1 maps, BORDER_CONSTANT 0
                        - the EuRoC rectification of /root/reference/examples/stereo/euroc.cc:118-133,176-177.
                        The arithmetic is OpenCV's (third-party, opencv 4.x imgproc/src/imgwarp.cpp): map
                        coordinates are rounded to 1/32 pixel (cvRound(v * 32), half to even), the four taps are
                        weighted with 15-bit integer coefficients (32-fx)(32-fy)*32 ... and the sum is rounded
                        with (s + 2^14) >> 15.  PINNED bit-for-bit against cv2.remap of this image
                        (tests/test_oracle_imgproc.py).
  * bgr_to_gray_u8      cv::cvtColor(BGR2GRAY) of the extractor's preprocess (src/SuperPoint.cc:387-388,770-771),
                        OpenCV 4.x 15-bit coefficients; PINNED against cv2 over all 2^24 colours.
  * undistort_points    cv::undistortPoints(raw, undist, K, D, noArray(), K) as called by
                        /root/reference/src/RgbdFrontEnd.cc:27-34: five fixed-point iterations of the
                        Brown-Conrady inverse in double precision, no FMA contraction.  PINNED against
                        cv2.undistortPoints.
  * rgbd_process        /root/reference/src/RgbdFrontEnd.cc:24-58 after the extraction: depth sampled at the
                        RAW keypoint (lround), uR = uL - bf / Z on the UNDISTORTED uL, has_depth iff 0 < Z < max.
                        PINNED against the reference's own RgbdFrontEnd.cc compiled in place (oracle/_ref,
                        tests/test_oracle_ref_frontend.py).
"""
from __future__ import annotations

import numpy as np

INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS
INTER_REMAP_COEF_BITS = 15


def fixed_point_maps(map_x: np.ndarray, map_y: np.ndarray):
    """cv::remap's float -> fixed-point map conversion: (ix, iy) int16 (saturated) and the 10-bit
    fractional index fy*32 + fx."""
    # cvRound of an out-of-range / NaN float is INT_MIN on x86 (cvtss2si)
    bad_x = ~np.isfinite(map_x) | (np.abs(map_x.astype(np.float64) * 32) >= 2 ** 31)
    bad_y = ~np.isfinite(map_y) | (np.abs(map_y.astype(np.float64) * 32) >= 2 ** 31)
    sx = np.rint(np.where(bad_x, 0, map_x).astype(np.float32) * np.float32(INTER_TAB_SIZE)).astype(np.int64)
    sy = np.rint(np.where(bad_y, 0, map_y).astype(np.float32) * np.float32(INTER_TAB_SIZE)).astype(np.int64)
    sx = np.where(bad_x, -2 ** 31, sx)
    sy = np.where(bad_y, -2 ** 31, sy)
    ix = np.clip(sx >> INTER_BITS, -32768, 32767).astype(np.int16)
    iy = np.clip(sy >> INTER_BITS, -32768, 32767).astype(np.int16)
    frac = ((sy & (INTER_TAB_SIZE - 1)) * INTER_TAB_SIZE + (sx & (INTER_TAB_SIZE - 1))).astype(np.uint16)
    return ix, iy, frac


def remap_linear_u8(src: np.ndarray, map_x: np.ndarray, map_y: np.ndarray) -> np.ndarray:
    """Bilinear remap of a u8 single-channel image, constant border 0."""
    assert src.dtype == np.uint8 and src.ndim == 2
    h, w = src.shape
    ix, iy, frac = fixed_point_maps(map_x, map_y)
    x0 = ix.astype(np.int64)
    y0 = iy.astype(np.int64)
    fx = (frac & 31).astype(np.int64)
    fy = (frac >> 5).astype(np.int64)
    w00 = (32 - fx) * (32 - fy) * 32
    w01 = fx * (32 - fy) * 32
    w10 = (32 - fx) * fy * 32
    w11 = fx * fy * 32

    def tap(y, x):
        ok = (x >= 0) & (x < w) & (y >= 0) & (y < h)
        v = src[np.clip(y, 0, h - 1), np.clip(x, 0, w - 1)].astype(np.int64)
        return np.where(ok, v, 0)

    s = tap(y0, x0) * w00 + tap(y0, x0 + 1) * w01 + tap(y0 + 1, x0) * w10 + tap(y0 + 1, x0 + 1) * w11
    out = (s + (1 << (INTER_REMAP_COEF_BITS - 1))) >> INTER_REMAP_COEF_BITS
    return np.clip(out, 0, 255).astype(np.uint8)


def undistort_points(xy: np.ndarray, fx: float, fy: float, cx: float, cy: float, dist) -> np.ndarray:
    """cv::undistortPoints(src, dst, K, D, noArray(), K): float32 [n, 2] in and out."""
    k = np.zeros(14, np.float64)
    d = np.asarray(dist, np.float64).ravel()
    k[: d.size] = d
    ifx, ify = 1.0 / np.float64(fx), 1.0 / np.float64(fy)
    out = np.empty((len(xy), 2), np.float32)
    for i, (u, v) in enumerate(np.asarray(xy, np.float32).astype(np.float64)):
        x = (u - cx) * ifx
        y = (v - cy) * ify
        x0, y0 = x, y
        for _ in range(5):
            r2 = x * x + y * y
            icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2)
            if icdist < 0:
                x, y = (u - cx) * ifx, (v - cy) * ify
                break
            dx = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2
            dy = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2
            x = (x0 - dx) * icdist
            y = (y0 - dy) * icdist
        xx = np.float64(fx) * x + 0.0 * y + np.float64(cx)
        yy = 0.0 * x + np.float64(fy) * y + np.float64(cy)
        ww = 1.0 / (0.0 * x + 0.0 * y + 1.0)
        out[i] = (np.float32(xx * ww), np.float32(yy * ww))
    return out


def bgr_to_gray_u8(bgr: np.ndarray) -> np.ndarray:
    """cv::cvtColor(img, g, cv::COLOR_BGR2GRAY) on u8 (src/SuperPoint.cc:387-388,770-771): OpenCV 4.x's 15-bit fixed
    point (B*3735 + G*19235 + R*9798 + 2^14) >> 15.  PINNED against cv2 on all 2^24 colours."""
    p = np.asarray(bgr, np.uint8).astype(np.int64)
    return ((p[..., 0] * 3735 + p[..., 1] * 19235 + p[..., 2] * 9798 + (1 << 14)) >> 15).astype(np.uint8)


def _lround(v: np.ndarray) -> np.ndarray:
    """std::lround: half away from zero."""
    v = v.astype(np.float64)
    return (np.sign(v) * np.floor(np.abs(v) + 0.5)).astype(np.int64)


def rgbd_process(xy: np.ndarray, depth: np.ndarray, fx, fy, cx, cy, dist, bf: float, depth_factor: float,
                 max_depth: float):
    """RgbdFrontEnd::process after extract: (undistorted xy float32 [n,2], stereo float64 [n,3], has_depth)."""
    xy = np.asarray(xy, np.float32).reshape(-1, 2)
    n = len(xy)
    has_dist = dist is not None and np.count_nonzero(np.asarray(dist)) > 0
    und = undistort_points(xy, fx, fy, cx, cy, dist) if (has_dist and n) else xy.copy()
    stereo = np.empty((n, 3), np.float64)
    has = np.zeros(n, np.int8)
    u = _lround(xy[:, 0])
    v = _lround(xy[:, 1])
    for i in range(n):
        z = 0.0
        if 0 <= u[i] < depth.shape[1] and 0 <= v[i] < depth.shape[0]:
            if depth.dtype == np.uint16:
                z = float(depth[v[i], u[i]]) / depth_factor
            elif depth.dtype == np.float32:
                z = float(np.float64(depth[v[i], u[i]]) / depth_factor)
        ul, vv = np.float64(und[i, 0]), np.float64(und[i, 1])
        if 0.0 < z < max_depth:
            stereo[i] = (ul, ul - bf / z, vv)
            has[i] = 1
        else:
            stereo[i] = (ul, np.nan, vv)
    return und, stereo, has
