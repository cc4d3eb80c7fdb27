"""The REFERENCE's own LightGlue host logic (/root/reference/src/LightGlue.cc compiled in place into
oracle/_ref/libref_nethost.so, TensorRT reduced to never-inferring stand-ins) against the restatements the product is
held to: what the engine is fed (LightGlue::prepare_inputs / store_keypoints, :228-283 - keypoints
normalised in float, descriptors converted to the fp16 binding) and how its outputs become cv::DMatch (row a12:
postprocess_outputs, :326-363)."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import frontend as ofe
from oracle import lightglue as olg

LIB = os.path.join(ROOT, "oracle", "_ref", "libref_nethost.so")
pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libref_nethost.so not built")
fp, ip, vp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_void_p


@pytest.fixture(scope="module")
def lib():
    lib = C.CDLL(LIB)
    lib.ref_lg_prepare_inputs.restype = C.c_int
    lib.ref_lg_prepare_inputs.argtypes = [C.c_int, C.c_int, fp, C.c_int, fp, fp, C.c_int, fp, C.c_int, C.c_int, vp, vp, vp, vp]
    lib.ref_lg_normalize_keypoints.argtypes = [C.c_int, C.c_int, fp, C.c_int, fp]
    lib.ref_lg_postprocess.restype = C.c_int
    lib.ref_lg_postprocess.argtypes = [ip, vp, C.c_int, C.c_int, ip, ip, fp]
    return lib


@pytest.mark.parametrize("w,h", [(640, 480), (1241, 376), (752, 480), (1280, 720), (375, 1242)])
def test_engine_inputs_keypoints_and_fp16_descriptors(lib, w, h):
    rng = np.random.default_rng(w + h)
    n0, n1 = 333, 1024
    xy0 = (rng.uniform(0, 1, (n0, 2)) * [w, h]).astype(np.float32)
    xy1 = (rng.integers(0, 8 * max(w, h), (n1, 2)) / 8.0).astype(np.float32)       # SuperPoint pixels scaled by W/W'
    xy1[:4] = [[0, 0], [w - 1, h - 1], [w / 2, h / 2], [w, h]]
    d0 = rng.normal(size=(n0, 256)).astype(np.float32)
    d1 = (rng.normal(size=(n1, 256)) * 10.0 ** rng.uniform(-9, 5, (n1, 1))).astype(np.float32)  # subnormal .. overflow
    k0, k1 = np.zeros((n0, 2), np.float32), np.zeros((n1, 2), np.float32)
    h0, h1 = np.zeros((n0, 256), np.float16), np.zeros((n1, 256), np.float16)
    # the reference's engine: kpts fp32, descriptors fp16 (scripts/rebuild_engines.sh:111-120)
    assert lib.ref_lg_prepare_inputs(w, h, xy0.ctypes.data_as(fp), n0, d0.ctypes.data_as(fp), xy1.ctypes.data_as(fp), n1,
                                     d1.ctypes.data_as(fp), 0, 1, k0.ctypes.data, k1.ctypes.data, h0.ctypes.data,
                                     h1.ctypes.data) == 1
    assert np.array_equal(k0, olg.normalize_keypoints(xy0, w, h)) and np.array_equal(k1, olg.normalize_keypoints(xy1, w, h))
    with np.errstate(over="ignore"):
        assert np.array_equal(h0.view(np.uint16), d0.astype(np.float16).view(np.uint16))   # round-to-nearest-even
        assert np.array_equal(h1.view(np.uint16), d1.astype(np.float16).view(np.uint16))
    assert np.isinf(h1).any() and (h1 == 0).any()
    # the float-only helper computes the same numbers
    k = np.zeros_like(k1)
    lib.ref_lg_normalize_keypoints(w, h, xy1.ctypes.data_as(fp), n1, k.ctypes.data_as(fp))
    assert np.array_equal(k, k1)
    # fp32 descriptor binding (an engine built without --fp16): passed through unchanged
    f0, f1 = np.zeros_like(d0), np.zeros_like(d1)
    assert lib.ref_lg_prepare_inputs(w, h, xy0.ctypes.data_as(fp), n0, d0.ctypes.data_as(fp), xy1.ctypes.data_as(fp), n1,
                                     d1.ctypes.data_as(fp), 0, 0, k0.ctypes.data, k1.ctypes.data, f0.ctypes.data,
                                     f1.ctypes.data) == 1
    assert np.array_equal(f0, d0) and np.array_equal(f1, d1)


@pytest.mark.parametrize("seed", range(4))
def test_outputs_to_dmatches(lib, seed):
    rng = np.random.default_rng(seed)
    n0 = [1, 64, 1024, 2048][seed]
    m0 = rng.integers(-1, 1500, n0).astype(np.int32)
    m0[rng.random(n0) < 0.4] = -1
    s0 = rng.random(n0).astype(np.float32)
    q, t, d = (np.zeros(n0, np.int32), np.zeros(n0, np.int32), np.zeros(n0, np.float32))
    n = lib.ref_lg_postprocess(m0.ctypes.data_as(ip), s0.ctypes.data, 0, n0, q.ctypes.data_as(ip), t.ctypes.data_as(ip),
                               d.ctypes.data_as(fp))
    eq, et, ed = ofe.dmatches(m0, s0)
    assert n == len(eq) and np.array_equal(q[:n], eq) and np.array_equal(t[:n], et) and np.array_equal(d[:n], ed)
    # an fp16 score binding is widened before the subtraction; no score binding -> distance 0
    s16 = s0.astype(np.float16)
    n = lib.ref_lg_postprocess(m0.ctypes.data_as(ip), s16.ctypes.data, 1, n0, q.ctypes.data_as(ip), t.ctypes.data_as(ip),
                               d.ctypes.data_as(fp))
    eq, et, ed = ofe.dmatches(m0, s16.astype(np.float32))
    assert n == len(eq) and np.array_equal(d[:n], ed)
    n = lib.ref_lg_postprocess(m0.ctypes.data_as(ip), None, 2, n0, q.ctypes.data_as(ip), t.ctypes.data_as(ip),
                               d.ctypes.data_as(fp))
    assert n == len(eq) and np.all(d[:n] == 0)


def test_no_keypoints_is_an_empty_result_not_an_error(lib):
    m0, s0 = np.zeros(1, np.int32), np.zeros(1, np.float32)
    q = np.zeros(1, np.int32)
    d = np.zeros(1, np.float32)
    assert lib.ref_lg_postprocess(m0.ctypes.data_as(ip), s0.ctypes.data, 0, 0, q.ctypes.data_as(ip), q.ctypes.data_as(ip),
                                  d.ctypes.data_as(fp)) == 0
