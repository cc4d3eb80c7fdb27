"""The REFERENCE's own SuperPoint::select_and_gather (/root/reference/src/SuperPoint.cc:681-750, compiled in place by
oracle/Makefile into oracle/_ref/libref_nethost.so; TensorRT reduced to never-called stand-ins) against the
restatement oracle/superpoint.py::select_keypoints that the CUDA nms / select kernels are held to bit for bit:
border strip, `float score > double threshold`, std::sort with std::greater on (score, (h, w)) - ties by row then column,
descending -, top-K, float scale factors for a resized score map.  Without a GPU the function stops at its pool check
after the keypoints are written (the slots are null), which is all the host half needs; the gather half runs in
tests/test_gpu_zz_ref_gather.py."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from oracle import superpoint as osp

LIB = os.path.join(ROOT, "oracle", "_ref", "libref_nethost.so")
pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libref_nethost.so not built")
fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)


def bind():
    lib = C.CDLL(LIB)
    lib.ref_sp_new.restype = C.c_void_p
    lib.ref_sp_new.argtypes = [C.c_int, C.c_double, C.c_int, C.c_int, C.c_int]
    lib.ref_sp_delete.argtypes = [C.c_void_p]
    lib.ref_sp_select_and_gather.restype = C.c_int
    lib.ref_sp_select_and_gather.argtypes = [C.c_void_p, fp, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, fp, fp, fp,
                                             C.c_void_p, ip, ip]
    return lib


def ref_select(lib, scores, input_h, input_w, K, thr, rb, grid_dev=None, want_desc=False):
    scores = np.ascontiguousarray(scores, np.float32)
    sh, sw = scores.shape
    sp = lib.ref_sp_new(K, thr, rb, input_h, input_w)
    xy, resp, sa = np.zeros((K, 2), np.float32), np.zeros(K, np.float32), np.zeros((K, 2), np.float32)
    desc = np.zeros((K, 256), np.uint16) if want_desc else None
    ok, info = C.c_int(-1), np.zeros(4, np.int32)
    n = lib.ref_sp_select_and_gather(sp, scores.ctypes.data_as(fp), sh, sw, grid_dev, sh // 8, sw // 8,
                                     xy.ctypes.data_as(fp), resp.ctypes.data_as(fp), sa.ctypes.data_as(fp),
                                     desc.ctypes.data if want_desc else None, C.byref(ok), info.ctypes.data_as(ip))
    lib.ref_sp_delete(sp)
    return dict(n=n, xy=xy[:n], score=resp[:n], size_angle=sa[:n], ok=ok.value, info=info,
                desc=None if desc is None else desc[:n])


def check(lib, scores, input_h, input_w, K, thr=0.005, rb=4):
    got = ref_select(lib, scores, input_h, input_w, K, thr, rb)
    exp = osp.select_keypoints(scores, input_h, input_w, K, thr, rb, scores.shape[0] // 8, scores.shape[1] // 8)
    assert got["n"] == len(exp["xy"])
    assert np.array_equal(got["xy"], exp["xy"]) and np.array_equal(got["score"], exp["score"])   # bits and order
    assert np.all(got["size_angle"] == np.float32([1.0, -1.0]))
    assert got["info"][0] == got["n"] and got["info"][1] == 256       # DescriptorPool::make(count): count, dim
    if got["n"] == 0:
        assert got["ok"] == 1
    return got, exp


@pytest.fixture(scope="module")
def lib():
    return bind()


@pytest.mark.parametrize("K", [64, 256, 100000])
def test_reference_module_score_maps(lib, K):
    """Score maps produced by the reference's own torch module (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLDEN, "superpoint_ref_small.npz"))
    for i in range(2):
        got, _ = check(lib, g["scores"][i], 120, 160, K)
        assert got["n"] > 30
    g = np.load(os.path.join(GOLDEN, "superpoint_ref_odd.npz"))       # 99 x 131 image, 96 x 128 map: scale != 1
    got, exp = check(lib, g["scores"][0], 99, 131, K)
    assert got["n"] > 10 and np.any(got["xy"][:, 0] != exp["hw"][:, 1])


def test_c2_candidates_of_the_reference_module(lib):
    g = np.load(os.path.join(GOLDEN, "superpoint_ref_c2.npz"))
    sh, sw = (int(x) for x in g["score_shape"])
    for i in range(2):
        m = np.zeros((sh, sw), np.float32)
        m[g[f"hw{i}"][:, 0], g[f"hw{i}"][:, 1]] = g[f"score{i}"]
        got, _ = check(lib, m, 480, 640, 1024)
        assert got["n"] == 1024


@pytest.mark.parametrize("seed", range(6))
def test_ties_threshold_edges_and_borders(lib, seed):
    rng = np.random.default_rng(seed)
    sh, sw = 8 * int(rng.integers(3, 9)), 8 * int(rng.integers(3, 11))
    thr = [0.005, 0.005, 0.0, 0.25, 0.015, 0.005][seed]
    rb = [4, 0, 7, 4, 1, 12][seed]
    # few distinct levels -> long runs of exactly equal scores (order decided by row, then column, descending)
    levels = np.float32([0.0, 0.0, 0.0, 0.004, 0.3, 0.3000001, 0.75, np.float32(thr), np.nextafter(np.float32(thr), np.float32(1)),
                         np.nextafter(np.float32(thr), np.float32(-1))])
    m = levels[rng.integers(0, len(levels), (sh, sw))]
    for K in (1, 17, 300, sh * sw):
        check(lib, m, sh * 2 + 1, sw * 3 + 2, K, thr, rb)


def test_nothing_above_threshold_and_degenerate_borders(lib):
    check(lib, np.zeros((48, 64), np.float32), 48, 64, 128)
    check(lib, np.full((48, 64), 0.5, np.float32), 48, 64, 128, rb=24)    # border strip swallows every row
    got, _ = check(lib, np.full((16, 16), 0.5, np.float32), 16, 16, 1000, rb=0)
    assert got["n"] == 256 and tuple(got["xy"][0]) == (15.0, 15.0) and tuple(got["xy"][-1]) == (0.0, 0.0)
