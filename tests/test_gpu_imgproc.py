"""GPU parity tests for the image front door, through the C-ABI: device remap and the RGB-D
keypoint post-process must equal the oracle (itself pinned bit-for-bit to cv2) exactly."""
import numpy as np
import pytest

from conftest import SP_WEIGHTS
from oracle import imgproc as oip
from superslam_b200.synth import synth_pair

pytestmark = pytest.mark.gpu

FX, FY, CX, CY = 458.654, 457.296, 367.215, 248.375
DIST = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0])


def _rectify_maps(h, w):
    """Brown-Conrady forward model on a rotated, re-projected pixel grid (what initUndistortRectifyMap
    evaluates), in numpy so the test does not depend on cv2 being importable on the GPU box."""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    x = (xx - 367.45) / 435.2
    y = (yy - 252.2) / 435.2
    a = 0.0077
    x, y = x * np.cos(a) - y * np.sin(a), x * np.sin(a) + y * np.cos(a)
    r2 = x * x + y * y
    kr = 1 + DIST[0] * r2 + DIST[1] * r2 * r2
    mx = FX * (x * kr + 2 * DIST[2] * x * y + DIST[3] * (r2 + 2 * x * x)) + CX
    my = FY * (y * kr + DIST[2] * (r2 + 2 * y * y) + 2 * DIST[3] * x * y) + CY
    return mx.astype(np.float32), my.astype(np.float32)


def test_remap_matches_opencv_arithmetic_exactly():
    from superslam_b200 import frontend as fe

    h, w = 480, 752
    left, right = synth_pair(h, w, 400, 77)
    mx, my = _rectify_maps(h, w)
    r = fe.Rectifier(mx, my, (h, w), max_images=2)
    out = r.remap([left, right])
    for src, got in zip((left, right), out):
        exp = oip.remap_linear_u8(src, mx, my)
        assert np.array_equal(got, exp)          # integer arithmetic: bit-exact
    try:
        import cv2
    except ImportError:
        return
    assert np.array_equal(out[0], cv2.remap(left, mx, my, cv2.INTER_LINEAR))


def test_remap_out_of_range_maps_and_ragged_size():
    from superslam_b200 import frontend as fe

    rng = np.random.default_rng(5)
    src, _ = synth_pair(99, 131, 60, 3)           # odd source size; destination 60 x 82 (multiple of 4 pixels)
    mx = rng.uniform(-25, 131 + 25, (60, 82)).astype(np.float32)
    my = rng.uniform(-25, 99 + 25, (60, 82)).astype(np.float32)
    mx[0, :4] = [np.nan, np.inf, -np.inf, 1e12]   # cvRound of these is INT_MIN -> far outside -> 0
    r = fe.Rectifier(mx, my, src.shape, max_images=1)
    got = r.remap([src])[0]
    assert np.array_equal(got, oip.remap_linear_u8(src, mx, my))
    assert got[0, :4].tolist() == [0, 0, 0, 0]


def test_rectify_then_extract_chain():
    """EuRoC flow: remap both images, then extract_stereo on the rectified pair (euroc.cc:176-181)."""
    from superslam_b200 import frontend as fe

    h, w = 480, 752
    left, right = synth_pair(h, w, 400, 21)
    mx, my = _rectify_maps(h, w)
    rect = fe.Rectifier(mx, my, (h, w))
    sp = fe.SuperPoint(SP_WEIGHTS, 512)
    l2, r2 = rect.remap([left, right])
    L, R = sp.extract_stereo(l2, r2)
    L0, R0 = sp.extract_stereo(oip.remap_linear_u8(left, mx, my), oip.remap_linear_u8(right, mx, my))
    assert np.array_equal(L.keypoints, L0.keypoints) and np.array_equal(R.keypoints, R0.keypoints)
    assert len(L.keypoints) > 20      # bilinear resampling smooths the synthetic corners: fewer detections


@pytest.mark.parametrize("dist", [None, DIST, np.array([-0.2, 0.05, 0.001, -0.0005, 0.01, 0.02, -0.01, 0.003])])
@pytest.mark.parametrize("dtype", [np.uint16, np.float32])
def test_rgbd_postprocess_exact(dist, dtype):
    from superslam_b200 import frontend as fe

    rng = np.random.default_rng(11)
    h, w, n = 480, 640, 1000
    xy = rng.uniform([-2, -2], [w + 2, h + 2], (n, 2)).astype(np.float32)   # a few raw points fall outside
    xy[:8] = np.floor(xy[:8]) + 0.5                                         # lround ties: away from zero
    if dtype == np.uint16:
        depth = rng.integers(0, 60000, (h, w)).astype(np.uint16)
        depth[rng.random((h, w)) < 0.2] = 0
        factor = 5000.0
    else:
        depth = rng.uniform(0, 12, (h, w)).astype(np.float32)
        factor = 1.0
    front = fe.RgbdFrontEnd(None, FX, FY, CX, CY, baseline=0.08, depth_factor=factor, max_depth=8.0,
                            dist_coeffs=dist, max_keypoints=1024, max_shape=(h, w))
    oxy, stereo, has = front.postprocess(xy, depth)
    exy, est, eh = oip.rgbd_process(xy, depth, FX, FY, CX, CY, dist, FX * 0.08, factor, 8.0)
    assert np.array_equal(oxy, exy)                         # fp64 without contraction: bit-exact float32
    assert np.array_equal(has, eh)
    assert np.array_equal(stereo, est, equal_nan=True)
    assert 0 < has.sum() < n


def test_rgbd_frontend_process_on_extracted_keypoints():
    from superslam_b200 import frontend as fe

    h, w = 480, 640
    gray, _ = synth_pair(h, w, 400, 31)
    depth = np.full((h, w), 10000, np.uint16)               # Z = 2 m everywhere
    sp = fe.SuperPoint(SP_WEIGHTS, 1000)                    # examples/rgbd/TUM1.yaml: max_keypoints 1000
    front = fe.RgbdFrontEnd(sp, 517.3, 516.5, 318.6, 255.3, baseline=40.0 / 517.3, depth_factor=5000.0,
                            max_depth=8.0, dist_coeffs=[0.2624, -0.9531, -0.0054, 0.0026, 1.1633],
                            max_keypoints=1000, max_shape=(h, w))
    f = front.process(gray, depth, 1.5)
    raw = sp.extract(gray).keypoints
    exy, est, eh = oip.rgbd_process(raw, depth, 517.3, 516.5, 318.6, 255.3,
                                    [0.2624, -0.9531, -0.0054, 0.0026, 1.1633], 517.3 * (40.0 / 517.3), 5000.0, 8.0)
    assert f.timestamp == 1.5 and len(f.keypoints_left) == len(raw) > 100
    assert np.array_equal(f.keypoints_left, exy) and np.array_equal(f.has_depth, eh)
    assert np.array_equal(f.stereo, est, equal_nan=True)
    assert f.has_depth.all()


def test_pipeline_with_device_rectification_equals_host_remap(lg_weights, tmp_path):
    """C4 flow: raw EuRoC-size pairs in, rectification + SuperPoint x2 + LightGlue + post-filter on the device;
    identical to rectifying with the (cv2-pinned) oracle first."""
    from superslam_b200 import frontend as fe
    from superslam_b200.lightglue_weights import save_state_dict

    h, w, K = 480, 752, 512
    lgw = str(tmp_path / "lg.ssbw")
    save_state_dict(lg_weights, lgw)
    mxl, myl = _rectify_maps(h, w)
    mxr, myr = mxl + np.float32(0.37), myl - np.float32(0.21)      # a different map for the right camera
    rl, rr = fe.Rectifier(mxl, myl, (h, w)), fe.Rectifier(mxr, myr, (h, w))
    pipe = fe.FramePairPipeline(SP_WEIGHTS, lgw, K, w, h, max_pairs=2)
    raw = []
    for s in (5, 6):
        raw += list(synth_pair(h, w, 400, s))
    rect = [oip.remap_linear_u8(im, *(m)) for im, m in zip(raw, [(mxl, myl), (mxr, myr)] * 2)]
    ref = pipe.process(rect)
    ref = {k: v.copy() for k, v in ref.items()}
    pipe.set_rectifiers(rl, rr)
    for _ in range(3):                      # eager, capture, replay
        got = pipe.process(raw)
        for k in ("count", "xy", "matches0", "has_depth"):
            assert np.array_equal(got[k], ref[k]), k
    assert int(ref["count"].min()) > 5        # interpolation smooths the synthetic corners: few detections
    pipe.set_rectifiers(None, None)
    back = pipe.process(rect)
    assert np.array_equal(back["matches0"], ref["matches0"])


def test_device_remap_and_undistortion_against_cv2_golden():
    """Committed outputs of cv2 itself (tests/golden/make_golden_imgproc.py): the device must reproduce them."""
    import os

    from conftest import GOLDEN
    from superslam_b200 import frontend as fe

    g = np.load(os.path.join(GOLDEN, "imgproc_cv2.npz"))
    img = np.ascontiguousarray(g["image"])
    for mx, my, exp in ((g["map_x"], g["map_y"], g["remap"]), (g["map_x_wide"], g["map_y_wide"], g["remap_wide"])):
        assert np.array_equal(fe.Rectifier(mx, my, img.shape, max_images=1).remap([img])[0], exp)
    fx, fy, cx, cy = (float(v) for v in g["camera"])
    depth = np.zeros((img.shape[0], img.shape[1]), np.uint16)
    for dist, exp in ((g["dist5"], g["undist5"]), (g["dist8"], g["undist8"])):
        front = fe.RgbdFrontEnd(None, fx, fy, cx, cy, baseline=0.1, depth_factor=5000.0, max_depth=8.0,
                                dist_coeffs=dist, max_keypoints=512, max_shape=img.shape)
        oxy, _, has = front.postprocess(g["points"], depth)
        assert np.array_equal(oxy, exp)
        assert not has.any()
