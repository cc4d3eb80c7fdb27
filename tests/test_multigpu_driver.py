"""The C++ multi-device driver (superslam_b200/csrc/multigpu.cpp, ssb_mg_*) is written above the public
C-ABI only, so it links against a CPU double of ssb_fe_* (tests/fake_fe.cpp): round-robin sharding of pairs over
devices, the walk in steps of max_pairs_per_device with one step in flight ahead, the scatter of every worker's rows
into the caller's arrays, and error propagation - without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("mg") / "libmg_fake.so")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-Wall", "-Wextra", "-pthread",
                           os.path.join(ROOT, "superslam_b200", "csrc", "multigpu.cpp"),
                           os.path.join(ROOT, "tests", "fake_fe.cpp"), "-Wl,--no-undefined", "-o", out])
    L = C.CDLL(out)
    L.ssb_mg_last_error.restype = C.c_char_p
    L.ssb_mg_last_error.argtypes = [C.c_void_p]
    L.ssb_mg_destroy.argtypes = [C.c_void_p]
    L.ssb_mg_device_of_pair.argtypes = [C.c_void_p, C.c_int]
    return L


def _expected(images, K, devices):
    pairs = len(images) // 2
    sums = [int(im.astype(np.uint64).sum()) & 0xffffffff for im in images]
    k = np.arange(K)
    o = dict(count=np.array([s % (K + 1) for s in sums], np.int32),
             xy=np.zeros((2 * pairs, K, 2), np.float32), score=np.zeros((2 * pairs, K), np.float32),
             matches0=np.zeros((pairs, K), np.int32), mscores0=np.zeros((pairs, K), np.float32),
             stereo_ur=np.zeros((pairs, K), np.float32), has_depth=np.zeros((pairs, K), np.uint8))
    for i, s in enumerate(sums):
        o["xy"][i, :, 0] = np.float32(s % 1000) + k.astype(np.float32)
        o["xy"][i, :, 1] = i & 1
        o["score"][i] = np.float32(s) * np.float32(0.5) - k.astype(np.float32)
    for p in range(pairs):
        s = sums[2 * p]
        o["matches0"][p] = (s + k) % 97 - 1
        o["mscores0"][p] = (((s + 3 * k) % 11).astype(np.float32) / np.float32(11.0))
        o["stereo_ur"][p] = devices[p % len(devices)]
        o["has_depth"][p] = (s + k) & 1
    return o


def _run(L, images, K, devices, max_pairs, null_outputs=()):
    pairs = len(images) // 2
    h, w = images[0].shape
    mg = C.c_void_p()
    ids = (C.c_int * len(devices))(*devices)
    assert L.ssb_mg_create(b"sp", b"lg", K, C.c_double(0.005), 4, w, h, C.c_float(1.0), max_pairs, ids, len(devices),
                           C.byref(mg)) == 0
    ptrs = (C.POINTER(C.c_uint8) * len(images))(*[i.ctypes.data_as(C.POINTER(C.c_uint8)) for i in images])
    o = dict(count=np.full((2 * pairs,), -7, np.int32), xy=np.zeros((2 * pairs, K, 2), np.float32),
             score=np.zeros((2 * pairs, K), np.float32), matches0=np.zeros((pairs, K), np.int32),
             mscores0=np.zeros((pairs, K), np.float32), stereo_ur=np.zeros((pairs, K), np.float32),
             has_depth=np.zeros((pairs, K), np.uint8))
    args = [None if k in null_outputs else o[k].ctypes.data_as(C.c_void_p)
            for k in ("count", "xy", "score", "matches0", "mscores0", "stereo_ur", "has_depth")]
    st = L.ssb_mg_process(mg, ptrs, pairs, h, w, images[0].strides[0], *args)
    err = L.ssb_mg_last_error(mg).decode()
    dev_of = [L.ssb_mg_device_of_pair(mg, p) for p in range(pairs)]
    L.ssb_mg_destroy(mg)
    return st, o, err, dev_of


@pytest.mark.parametrize("pairs,devices,max_pairs", [(11, [0, 1, 2], 2), (8, [3, 1], 64), (2, [0, 1, 2, 3, 4, 5, 6, 7], 4),
                                                    (17, [5], 3), (1, [0, 1], 1)])
def test_sharding_walk_and_scatter(lib, pairs, devices, max_pairs):
    rng = np.random.default_rng(pairs)
    K = 13
    images = [rng.integers(0, 256, (6, 10), np.uint8) for _ in range(2 * pairs)]
    st, got, err, dev_of = _run(lib, images, K, devices, max_pairs)
    assert st == 0, err
    exp = _expected(images, K, devices)
    for k in exp:
        assert np.array_equal(got[k], exp[k]), k
    assert dev_of == [devices[p % len(devices)] for p in range(pairs)]      # pair p -> device p mod G


def test_strided_images_and_null_outputs(lib):
    rng = np.random.default_rng(5)
    big = [rng.integers(0, 256, (6, 16), np.uint8) for _ in range(10)]
    views = [b[:, :10] for b in big]                      # row stride 16, width 10
    st, got, err, _ = _run(lib, views, 7, [0, 1], 2, null_outputs=("xy", "mscores0"))
    assert st == 0, err
    exp = _expected([np.ascontiguousarray(v) for v in views], 7, [0, 1])
    for k in ("count", "score", "matches0", "stereo_ur", "has_depth"):
        assert np.array_equal(got[k], exp[k]), k
    assert not got["xy"].any() and not got["mscores0"].any()          # NULL outputs are left alone


def test_a_failing_device_fails_the_call_with_its_name(lib):
    rng = np.random.default_rng(9)
    images = [rng.integers(0, 256, (4, 8), np.uint8) for _ in range(12)]
    lib.fake_fe_fail_on_device(1)
    try:
        st, _, err, _ = _run(lib, images, 5, [0, 1, 2], 1)
    finally:
        lib.fake_fe_fail_on_device(-1)
    assert st == 2 and "device 1" in err and "injected failure" in err
    st, got, err, _ = _run(lib, images, 5, [0, 1, 2], 1)               # and the driver is usable again afterwards
    assert st == 0 and np.array_equal(got["count"], _expected(images, 5, [0, 1, 2])["count"])
