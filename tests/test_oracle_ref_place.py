"""The REFERENCE's own retrieval code (/root/reference/src/PlaceRecognizer.cc compiled in place into
oracle/_ref/libref_place.so) against the restatement in oracle/eigenplaces.py that the device index (ssb_ep_add /
ssb_ep_query, tests/test_gpu_eigenplaces.py) is held to: recency window, score gate, descending order, top-K, and the
vote streaks of TemporalConsistencyVoter.  Control flow is compared exactly; scores to 1e-6 (cv::norm / gemm arithmetic is
the cv::Mat stand-in's, see oracle/stubs_cv)."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import eigenplaces as oep

LIB = os.path.join(ROOT, "oracle", "_ref", "libref_place.so")
pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libref_place.so not built")
fp = C.POINTER(C.c_float)


@pytest.fixture(scope="module")
def lib():
    lib = C.CDLL(LIB)
    lib.ref_index_new.restype = C.c_void_p
    lib.ref_index_delete.argtypes = [C.c_void_p]
    lib.ref_index_size.argtypes = [C.c_void_p]
    lib.ref_index_add.argtypes = [C.c_void_p, C.c_size_t, fp, C.c_int, C.c_int]
    lib.ref_index_query.restype = C.c_int
    lib.ref_index_query.argtypes = [C.c_void_p, fp, C.c_int, C.c_size_t, C.c_int, C.c_float, C.POINTER(C.c_size_t), fp, C.c_int]
    lib.ref_voter_new.restype = C.c_void_p
    lib.ref_voter_new.argtypes = [C.c_int, C.c_size_t]
    lib.ref_voter_delete.argtypes = [C.c_void_p]
    lib.ref_voter_vote.restype = C.c_int
    lib.ref_voter_vote.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_float]
    return lib


def ref_query(lib, h, d, exclude, top_k, min_score, cap=4096):
    ids, sc = (C.c_size_t * cap)(), np.zeros(cap, np.float32)
    d = np.ascontiguousarray(d, np.float32)
    n = lib.ref_index_query(h, d.ctypes.data_as(fp), d.size, exclude, top_k, min_score, ids, sc.ctypes.data_as(fp), cap)
    return [int(ids[i]) for i in range(n)], sc[:n].copy()


@pytest.mark.parametrize("seed,dim", [(0, 512), (1, 512), (2, 64)])
def test_index_against_the_reference_code(lib, seed, dim):
    rng = np.random.default_rng(seed)
    h, mine = lib.ref_index_new(), oep.CosineDescriptorIndex()
    base = rng.normal(size=dim).astype(np.float32)
    assert ref_query(lib, h, base, 0, 5, 0.0)[0] == [] and mine.query(base, 0, 5, 0.0) == []      # empty database
    for k in range(120):
        d = (base * rng.uniform(0.2, 3.0) + rng.normal(size=dim) * rng.uniform(0.05, 4.0)).astype(np.float32)  # not unit length
        kid = 1000 + 7 * k
        lib.ref_index_add(h, kid, np.ascontiguousarray(d).ctypes.data_as(fp), dim, k % 2)
        mine.add(kid, d)
        if k in (0, 1, 30, 119):
            assert lib.ref_index_size(h) == mine.size() == k + 1
            for exclude, top_k, min_score in [(0, 5, 0.75), (0, 0, -1.0), (10, 3, 0.2), (k, 0, -1.0), (k + 1, 5, -1.0),
                                              (k + 50, 5, -1.0), (5, 1000, 0.5), (0, 1, 0.999)]:
                q = (base + rng.normal(size=dim) * 0.3).astype(np.float32)
                ids, sc = ref_query(lib, h, q, exclude, top_k, min_score)
                exp = mine.query(q, exclude, top_k, min_score)
                assert ids == [e[0] for e in exp], (k, exclude, top_k, min_score)
                assert np.allclose(sc, [e[1] for e in exp], atol=1e-6) and np.all(np.diff(sc) <= 0)
                assert all(s >= min_score - 1e-6 for s in sc) and (top_k <= 0 or len(ids) <= top_k)
    # the reference's unit test (tests/test_place_recognizer.cc): a query equal to a stored row scores 1 and comes first
    probe = (base * 2.0).astype(np.float32)
    lib.ref_index_add(h, 5, probe.ctypes.data_as(fp), dim, 0)
    mine.add(5, probe)
    ids, sc = ref_query(lib, h, probe, 0, 1, 0.9)
    assert ids == [5] == [mine.query(probe, 0, 1, 0.9)[0][0]] and abs(sc[0] - 1.0) < 1e-6
    assert ref_query(lib, h, probe, 1, 1, 0.9999)[0] == []                 # ... unless it is inside the recency window
    lib.ref_index_delete(h)


@pytest.mark.parametrize("required,tol", [(1, 0), (3, 5), (2, 0), (4, 100)])
def test_voter_against_the_reference_code(lib, required, tol):
    rng = np.random.default_rng(required * 31 + tol)
    h, mine = lib.ref_voter_new(required, tol), oep.TemporalConsistencyVoter(required, tol)
    kid = 50
    for _ in range(400):
        r = rng.random()
        if r < 0.15:
            assert lib.ref_voter_vote(h, 0, 0, 0.0) == int(mine.vote(None)) == 0
            continue
        kid = max(0, kid + int(rng.integers(-tol - 2, tol + 3))) if r < 0.9 else int(rng.integers(0, 10000))
        assert lib.ref_voter_vote(h, 1, kid, 0.9) == int(mine.vote((kid, 0.9)))
    lib.ref_voter_delete(h)


# ---- EigenPlaces::preprocess (src/EigenPlaces.cc:123-143), the reference's own code, cv::resize served by cv2 --------------
NETHOST = os.path.join(ROOT, "oracle", "_ref", "libref_nethost.so")
RESIZE_FN = C.CFUNCTYPE(None, C.POINTER(C.c_ubyte), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_ubyte), C.c_int, C.c_int)


@pytest.mark.skipif(not os.path.exists(NETHOST), reason="oracle/_ref/libref_nethost.so not built")
@pytest.mark.parametrize("shape,net", [((480, 752), (512, 512)), ((480, 640, 3), (512, 512)), ((480, 640), (640, 480)),
                                       ((1024, 1024, 3), (512, 512)), ((99, 131, 3), (64, 96)), ((376, 1241), (224, 256))])
def test_eigenplaces_preprocess_against_the_reference_code(shape, net):
    """gray / BGR -> RGB, resize (incl. the exact 2x decimation and the identity size), /255, ImageNet normalisation, CHW:
    oracle/eigenplaces.py::preprocess - what the device preprocess kernel is compared with - equals the reference's
    function bit for bit."""
    import cv2

    lib = C.CDLL(NETHOST)
    lib.ref_ep_preprocess.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, RESIZE_FN, fp]
    calls = []

    @RESIZE_FN
    def resize(src, sh, sw, ch, step, dst, dh, dw):
        a = np.ctypeslib.as_array(src, shape=(sh, step))[:, :sw * ch].reshape(sh, sw, ch)
        out = cv2.resize(np.ascontiguousarray(a), (dw, dh))                 # INTER_LINEAR, as cv::resize defaults to
        np.ctypeslib.as_array(dst, shape=(dh, dw * ch))[:] = out.reshape(dh, dw * ch)
        calls.append((sh, sw, ch, dh, dw))

    rng = np.random.default_rng(sum(shape))
    ch = 1 if len(shape) == 2 else 3
    buf = rng.integers(0, 256, (shape[0], shape[1] * ch + 13), dtype=np.uint8)     # padded rows: cv::Mat step
    img = buf[:, :shape[1] * ch].reshape(shape)
    in_w, in_h = net
    got = np.zeros((3, in_h, in_w), np.float32)
    lib.ref_ep_preprocess(img.ctypes.data, shape[0], shape[1], ch, buf.strides[0], in_w, in_h, resize, got.ctypes.data_as(fp))
    assert calls == [(shape[0], shape[1], 3, in_h, in_w)]                   # resized once, after the colour conversion
    exp = oep.preprocess(img, in_w, in_h)
    assert exp.shape == got.shape and np.array_equal(exp.view(np.uint32), got.view(np.uint32))
