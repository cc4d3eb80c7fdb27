"""Oracle pinning for the image front door: oracle/imgproc.py against OpenCV itself
(cv2 of this image runs the same C++ the reference links: cv::remap, cv::undistortPoints)."""
import numpy as np
import pytest

from oracle import imgproc as ip
from superslam_b200.synth import synth_pair

try:
    import cv2
except ImportError:   # the committed golden vectors below still pin the oracle
    cv2 = None
needs_cv2 = pytest.mark.skipif(cv2 is None, reason="cv2 not importable")

K = np.array([[458.654, 0, 367.215], [0, 457.296, 248.375], [0, 0, 1]])
D = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0])   # EuRoC cam0
P = np.array([[435.2046959714599, 0, 367.4517211914062], [0, 435.2046959714599, 252.2008514404297], [0, 0, 1]])


def euroc_maps(h=480, w=752):
    R = cv2.Rodrigues(np.array([0.0077, -0.0049, 0.0016]))[0]
    return cv2.initUndistortRectifyMap(K, D, R, P, (w, h), cv2.CV_32F)


@needs_cv2
def test_remap_rectification_bit_exact():
    img, _ = synth_pair(480, 752, 300, 5)
    m1, m2 = euroc_maps()
    assert np.array_equal(cv2.remap(img, m1, m2, cv2.INTER_LINEAR), ip.remap_linear_u8(img, m1, m2))


@needs_cv2
def test_remap_border_and_ties_bit_exact():
    rng = np.random.default_rng(3)
    img, _ = synth_pair(120, 160, 60, 9)
    h, w = img.shape
    # coordinates well outside the image on every side (constant border 0) ...
    mx = rng.uniform(-20, w + 20, (h, w)).astype(np.float32)
    my = rng.uniform(-20, h + 20, (h, w)).astype(np.float32)
    assert np.array_equal(cv2.remap(img, mx, my, cv2.INTER_LINEAR), ip.remap_linear_u8(img, mx, my))
    # ... and exact 1/64 offsets: cvRound(v * 32) ties go to even
    mx = (np.arange(w)[None, :] + np.zeros((h, 1)) + 1 / 64).astype(np.float32)
    my = (np.arange(h)[:, None] + np.zeros((1, w)) + 3 / 64).astype(np.float32)
    assert np.array_equal(cv2.remap(img, mx, my, cv2.INTER_LINEAR), ip.remap_linear_u8(img, mx, my))


@needs_cv2
@pytest.mark.parametrize("dist", [D, np.array([-0.2, 0.05, 0.001, -0.0005, 0.01, 0.02, -0.01, 0.003])])
def test_undistort_points_bit_exact(dist):
    rng = np.random.default_rng(1)
    pts = rng.uniform([0, 0], [752, 480], (1500, 2)).astype(np.float32)
    ref = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, dist, None, K).reshape(-1, 2)
    got = ip.undistort_points(pts, K[0, 0], K[1, 1], K[0, 2], K[1, 2], dist)
    assert np.array_equal(ref, got)


def test_rgbd_process_semantics():
    # src/RgbdFrontEnd.cc:24-58: depth at the RAW pixel, uR from the UNDISTORTED uL, Z window (0, max_depth)
    xy = np.array([[10.4, 20.6], [100.5, 50.5], [-3.0, 5.0], [30.0, 40.0]], np.float32)
    depth = np.zeros((120, 160), np.uint16)
    depth[21, 10] = 5000      # lround(10.4, 20.6) = (10, 21): Z = 1.0
    depth[51, 101] = 50000    # lround(.5) goes away from zero: (101, 51); Z = 10 -> beyond max_depth 8
    depth[40, 30] = 0         # no measurement
    und, stereo, has = ip.rgbd_process(xy, depth, 500.0, 500.0, 80.0, 60.0, None, bf=40.0, depth_factor=5000.0,
                                       max_depth=8.0)
    assert np.array_equal(und, xy)                      # no distortion: keypoints unchanged
    assert has.tolist() == [1, 0, 0, 0]
    assert stereo[0].tolist() == [float(xy[0, 0]), float(xy[0, 0]) - 40.0, float(xy[0, 1])]
    assert np.isnan(stereo[1:, 1]).all()


def test_oracle_against_committed_cv2_golden():
    """The same pin without OpenCV at hand: tests/golden/imgproc_cv2.npz holds cv2's own outputs
    (tests/golden/make_golden_imgproc.py) and travels to the GPU box."""
    import os

    from conftest import GOLDEN

    g = np.load(os.path.join(GOLDEN, "imgproc_cv2.npz"))
    assert np.array_equal(ip.remap_linear_u8(g["image"], g["map_x"], g["map_y"]), g["remap"])
    assert np.array_equal(ip.remap_linear_u8(g["image"], g["map_x_wide"], g["map_y_wide"]), g["remap_wide"])
    assert (g["remap_wide"] == 0).sum() > 100            # the wide map really leaves the source image
    fx, fy, cx, cy = g["camera"]
    assert np.array_equal(ip.undistort_points(g["points"], fx, fy, cx, cy, g["dist5"]), g["undist5"])
    assert np.array_equal(ip.undistort_points(g["points"], fx, fy, cx, cy, g["dist8"]), g["undist8"])


def test_remap_properties():
    """Size-independent properties of the fixed-point remap: identity maps reproduce the image, integer shifts move
    it (zero-filled where the source ends), half-pixel maps average neighbours with OpenCV's rounding."""
    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, (37, 53)).astype(np.uint8)
    h, w = img.shape
    xx, yy = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    assert np.array_equal(ip.remap_linear_u8(img, xx, yy), img)
    sh = ip.remap_linear_u8(img, xx + 5, yy - 3)            # dst(y, x) = src(y - 3, x + 5)
    exp = np.zeros_like(img)
    exp[3:, : w - 5] = img[: h - 3, 5:]
    assert np.array_equal(sh, exp)
    half = ip.remap_linear_u8(img, xx + 0.5, yy)            # (a + b) / 2 with (s + 2^14) >> 15 rounding: ties go up
    a, b = img[:, :-1].astype(np.int64), img[:, 1:].astype(np.int64)
    assert np.array_equal(half[:, :-1], ((a + b + 1) >> 1).astype(np.uint8))
    assert np.array_equal(half[:, -1], ((img[:, -1].astype(np.int64) + 1) >> 1).astype(np.uint8))   # right tap is border 0


def test_undistort_properties():
    """Zero distortion is the identity (to float rounding of the round trip), the principal point is a fixed point,
    and undistorting a forward-distorted point recovers it as far as five iterations go."""
    fx, fy, cx, cy = 458.654, 457.296, 367.215, 248.375
    rng = np.random.default_rng(4)
    pts = rng.uniform([0, 0], [752, 480], (500, 2)).astype(np.float32)
    same = ip.undistort_points(pts, fx, fy, cx, cy, np.zeros(5))
    assert np.abs(same - pts).max() < 1e-4
    c = ip.undistort_points(np.array([[cx, cy]], np.float32), fx, fy, cx, cy, D)
    assert np.abs(c - np.array([[cx, cy]], np.float32)).max() < 1e-4
    # forward Brown-Conrady model, then the inverse
    x, y = (pts[:, 0].astype(np.float64) - cx) / fx, (pts[:, 1].astype(np.float64) - cy) / fy
    r2 = x * x + y * y
    kr = 1 + D[0] * r2 + D[1] * r2 * r2 + D[4] * r2 ** 3
    xd = x * kr + 2 * D[2] * x * y + D[3] * (r2 + 2 * x * x)
    yd = y * kr + D[2] * (r2 + 2 * y * y) + 2 * D[3] * x * y
    dist = np.stack([xd * fx + cx, yd * fy + cy], 1).astype(np.float32)
    back = ip.undistort_points(dist, fx, fy, cx, cy, D)
    err = np.abs(back - pts).max(axis=1)
    assert err[r2 < 0.25].max() < 2e-2      # five fixed-point iterations converge fast near the centre ...
    assert err.max() < 0.5                  # ... and slowly in the corners (cv::undistortPoints' own behaviour)


def test_library_map_conversion_matches_opencv_fixed_point():
    """Host logic of the product, no GPU: ssb_rect_convert_maps (what ssb_rect_create uploads) equals the oracle's
    restatement of cv::remap's map conversion - ties to even, saturation, NaN / infinity -> INT_MIN."""
    import ctypes as C

    from superslam_b200 import _lib

    rng = np.random.default_rng(9)
    mx = rng.uniform(-40000, 40000, 20000).astype(np.float32)      # beyond the int16 range: saturation
    my = rng.uniform(-50, 600, 20000).astype(np.float32)
    mx[:6] = [np.nan, np.inf, -np.inf, 1e12, 0.015625, 0.046875]   # 1/64 and 3/64: cvRound ties
    my[:6] = [3.0, 4.0, 5.0, -1e12, 0.015625, 0.046875]
    xy = np.zeros(len(mx), np.uint32)
    fr = np.zeros(len(mx), np.uint16)
    _lib.check(_lib.load().ssb_rect_convert_maps(mx.ctypes.data_as(C.POINTER(C.c_float)),
                                                 my.ctypes.data_as(C.POINTER(C.c_float)), len(mx),
                                                 xy.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                 fr.ctypes.data_as(C.POINTER(C.c_uint16))))
    ix, iy, frac = ip.fixed_point_maps(mx, my)
    assert np.array_equal((xy & 0xffff).astype(np.uint16).view(np.int16), ix)
    assert np.array_equal((xy >> 16).astype(np.uint16).view(np.int16), iy)
    assert np.array_equal(fr, frac)
    assert fr[4] == 0 and fr[5] == (2 * 32 + 2)                    # 0.5 -> 0 and 1.5 -> 2: round half to even


@needs_cv2
def test_bgr_to_gray_matches_opencv_on_every_colour():
    """cv::cvtColor(BGR2GRAY) as the extractor's preprocess applies it (src/SuperPoint.cc:387-388,770-771): all 2^24
    colours against cv2 (OpenCV 4.x uses 15-bit coefficients; the older 14-bit set differs on ~0.24 % of pixels)."""
    b, g, r = np.meshgrid(np.arange(256), np.arange(256), np.arange(256), indexing="ij")
    cols = np.stack([b, g, r], -1).reshape(4096, 4096, 3).astype(np.uint8)
    assert np.array_equal(ip.bgr_to_gray_u8(cols), cv2.cvtColor(cols, cv2.COLOR_BGR2GRAY))
    old = ((cols.astype(np.int64) @ np.array([1868, 9617, 4899]) + 8192) >> 14).astype(np.uint8)
    assert (old != cv2.cvtColor(cols, cv2.COLOR_BGR2GRAY)).mean() > 1e-3     # the two coefficient sets do differ


def test_bgr_to_gray_against_committed_cv2_golden():
    import os

    from conftest import GOLDEN

    g = np.load(os.path.join(GOLDEN, "imgproc_cv2.npz"))
    assert np.array_equal(ip.bgr_to_gray_u8(g["bgr"]), g["gray"])
    assert len(np.unique(g["bgr"].reshape(-1, 3), axis=0)) > 1000      # a real colour image, not B = G = R
