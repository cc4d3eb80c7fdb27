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
