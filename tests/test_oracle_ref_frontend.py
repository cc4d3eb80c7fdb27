"""oracle/_ref: the reference's OWN StereoFrontEnd::process and StereoFrame::backproject (src/StereoFrontEnd.cc,
src/StereoFrame.cc, compiled in place by oracle/Makefile behind mock extractor / matcher objects) pin the restated
post-filter of oracle/frontend.py - the arithmetic the device kernel stereo_postfilter_kernel reproduces - including
its float-vs-double compare semantics at the thresholds."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle.frontend import stereo_postfilter

LIB = os.path.join(ROOT, "oracle", "_ref", "libref_frontend.so")
pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref not built (needs /root/reference: build())")

_f = C.POINTER(C.c_float)
_i = C.POINTER(C.c_int)
_d = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def ref():
    lib = C.CDLL(LIB)
    lib.ref_stereo_frontend_process.argtypes = [_f, C.c_int, _f, C.c_int, _i, _i, C.c_int, C.c_float, _d, C.c_char_p]
    lib.ref_stereo_frame_backproject.argtypes = [_d, _d, _d, _d, _d]
    return lib


def _run(ref, xl, xr, q, t, min_disp=1.0):
    xl, xr = np.ascontiguousarray(xl, np.float32), np.ascontiguousarray(xr, np.float32)
    q, t = np.ascontiguousarray(q, np.int32), np.ascontiguousarray(t, np.int32)
    st = np.zeros((len(xl), 3), np.float64)
    has = C.create_string_buffer(max(1, len(xl)))
    ref.ref_stereo_frontend_process(xl.ctypes.data_as(_f), len(xl), xr.ctypes.data_as(_f), len(xr), q.ctypes.data_as(_i),
                                    t.ctypes.data_as(_i), len(q), C.c_float(min_disp), st.ctypes.data_as(_d), has)
    return st, np.frombuffer(has.raw[: len(xl)], np.int8).copy()


def test_reference_unit_test_cases(ref):
    """tests/test_stereo_frontend.cc:49-73 on the real code: a 12-px disparity on the same row is kept; a negative
    disparity and a 5-px row offset are rejected and stay (uL, NaN, v)."""
    xl = np.array([[100, 50], [200, 80], [300, 120]], np.float32)
    xr = np.array([[88, 50], [210, 80], [290, 125]], np.float32)
    st, has = _run(ref, xl, xr, [0, 1, 2], [0, 1, 2])
    assert has.tolist() == [1, 0, 0]
    assert st[0].tolist() == [100.0, 88.0, 50.0] and np.isnan(st[1, 1]) and np.isnan(st[2, 1])
    assert st[1, 0] == 200.0 and st[2, 2] == 120.0


def test_restated_postfilter_equals_the_reference_on_random_and_boundary_inputs(ref):
    rng = np.random.default_rng(3)
    for trial in range(30):
        nl, nr = int(rng.integers(1, 200)), int(rng.integers(1, 200))
        xl = rng.uniform(0, 640, (nl, 2)).astype(np.float32)
        xr = rng.uniform(0, 640, (nr, 2)).astype(np.float32)
        nm = int(rng.integers(0, 300))
        q = rng.integers(-2, nl + 2, nm).astype(np.int32)          # some indices out of range on purpose
        t = rng.integers(-2, nr + 2, nm).astype(np.int32)
        for k in range(min(nm, 40)):                                # force cases AT the thresholds
            i, j = int(q[k]), int(t[k])
            if 0 <= i < nl and 0 <= j < nr:
                base = np.float32(rng.uniform(10, 600))
                d = rng.choice([1.0, np.nextafter(np.float32(1.0), np.float32(0)), np.nextafter(np.float32(1.0), np.float32(2)),
                                0.99999, 12.0])
                xl[i, 0] = base + np.float32(d)
                xr[j, 0] = base
                dv = rng.choice([2.0, np.nextafter(np.float32(2.0), np.float32(3)), -2.0, 1.5, 2.0001])
                xr[j, 1] = xl[i, 1] - np.float32(dv)
        md = float(rng.choice([1.0, 0.5, 2.5]))
        st_ref, has_ref = _run(ref, xl, xr, q, t, md)
        st, has = stereo_postfilter(xl, xr, q, t, md)
        assert np.array_equal(has, has_ref), trial
        assert np.array_equal(st, st_ref, equal_nan=True), trial


def test_backproject_reference_vs_python_mirror(ref):
    """tests/test_stereo_frame.cc:10-23 numbers through the real StereoFrame::backproject and through the mirror."""
    from superslam_b200.frontend import StereoFrame

    fx = fy = 500.0
    cx, cy, b = 320.0, 240.0, 0.5
    a = 0.1
    R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    t = np.array([2.0, -1.0, 0.5])
    world = np.array([3.0, 0.5, 9.0])
    pc = R.T @ (world - t)
    uL = fx * pc[0] / pc[2] + cx
    z = np.array([uL, uL - fx * b / pc[2], fy * pc[1] / pc[2] + cy])
    out = np.zeros(3)
    ref.ref_stereo_frame_backproject(z.ctypes.data_as(_d), np.ascontiguousarray(R).ctypes.data_as(_d), t.ctypes.data_as(_d),
                                     np.array([fx, fy, cx, cy, b]).ctypes.data_as(_d), out.ctypes.data_as(_d))
    assert np.allclose(out, world, atol=1e-4)
    mirror = StereoFrame(0.0, None, None, z[None], np.array([1], np.int8), R, t).backproject(0, fx, fy, cx, cy, b)
    assert np.allclose(mirror, out, rtol=0, atol=1e-12)


@pytest.mark.parametrize("dtype", [np.uint16, np.float32])
@pytest.mark.parametrize("with_distortion", [False, True])
def test_restated_rgbd_process_equals_the_reference(ref, dtype, with_distortion):
    """The real RgbdFrontEnd::process (src/RgbdFrontEnd.cc:23-58, compiled in place) with its cv::undistortPoints call
    served by the real OpenCV (cv2) through a callback, against oracle/imgproc.py::rgbd_process: undistorted
    keypoints, depth sampled at lround(raw), uR = uL - bf / Z, the (0, max_depth) window - bit for bit."""
    from oracle import imgproc as oip

    cv2 = pytest.importorskip("cv2")
    CB = C.CFUNCTYPE(None, _f, C.c_int, _d, _d, C.c_int, _d, _f)

    def undistort(src, n, K, D, nd, P, dst):
        pts = np.ctypeslib.as_array(src, (n, 2)).copy()
        k = np.ctypeslib.as_array(K, (9,)).reshape(3, 3).copy()
        p = np.ctypeslib.as_array(P, (9,)).reshape(3, 3).copy()
        d = np.ctypeslib.as_array(D, (max(nd, 1),))[:nd].copy()
        out = cv2.undistortPoints(pts.reshape(-1, 1, 2), k, d, None, p).reshape(-1, 2).astype(np.float32)
        np.ctypeslib.as_array(dst, (n, 2))[:] = out

    cb = CB(undistort)
    ref.ref_rgbd_frontend_process.argtypes = [_f, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, _d, C.c_double, _d,
                                              C.c_int, C.c_double, C.c_double, CB, _f, _d, C.c_char_p]
    rng = np.random.default_rng(21)
    h, w, n = 120, 160, 400
    fx, fy, cx, cy, baseline = 130.0, 129.0, 80.5, 59.5, 0.08
    xy = rng.uniform([-2, -2], [w + 2, h + 2], (n, 2)).astype(np.float32)      # a few raw points fall outside the depth map
    xy[:10] = np.floor(xy[:10]) + np.float32(0.5)                              # lround ties: away from zero
    if dtype == np.uint16:
        depth = rng.integers(0, 60000, (h, w)).astype(np.uint16)
        depth[rng.random((h, w)) < 0.2] = 0
        factor = 5000.0
    else:
        depth = rng.uniform(0, 12, (h, w)).astype(np.float32)
        factor = 1.0
    depth[3, 7] = 40000 if dtype == np.uint16 else 8.0                         # Z == max_depth exactly: rejected (strict <)
    xy[10] = (7.2, 3.1)
    dist = np.array([-0.28, 0.07, 0.0002, 1.8e-05, 0.0]) if with_distortion else np.zeros(5)
    oxy = np.zeros((n, 2), np.float32)
    st = np.zeros((n, 3), np.float64)
    has = C.create_string_buffer(n)
    cam = np.array([fx, fy, cx, cy])
    dd = dist.copy()
    ref.ref_rgbd_frontend_process(xy.ctypes.data_as(_f), n, depth.ctypes.data_as(C.c_void_p), 0 if dtype == np.uint16 else 1,
                                  h, w, cam.ctypes.data_as(_d), baseline, dd.ctypes.data_as(_d), len(dd), factor, 8.0, cb,
                                  oxy.ctypes.data_as(_f), st.ctypes.data_as(_d), has)
    has_ref = np.frombuffer(has.raw[:n], np.int8)
    exy, est, eh = oip.rgbd_process(xy, depth, fx, fy, cx, cy, dist, fx * baseline, factor, 8.0)
    assert np.array_equal(oxy, exy)
    assert np.array_equal(has_ref, eh)
    assert np.array_equal(st, est, equal_nan=True)
    assert has_ref[10] == 0 and 0 < has_ref.sum() < n
    if with_distortion:
        assert np.abs(oxy - xy).max() > 0.05          # the undistortion really moved the keypoints
    else:
        assert np.array_equal(oxy, xy)                # countNonZero(dist) == 0: undistortPoints is not called
