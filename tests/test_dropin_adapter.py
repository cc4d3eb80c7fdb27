"""The C++ half of the drop-in boundary, EXECUTED: the reference's own StereoFrontEnd::process
(/root/reference/src/StereoFrontEnd.cc, compiled in place) runs on top of include/superslam_b200_adapter.hpp, plus the
other call shapes the reference makes through IFeatureMatcher (tracking match, descriptors_to_host, loop verification on
a cloned context) - oracle/dropin_harness.cpp.

On the CPU the C-ABI below the adapter is a test double with canned behaviour (oracle/fake_capi.cpp -> libdropin_fake.so)
so that what is checked is the adapter's own work: cv::Mat -> (pointer, stride, channels), cv::KeyPoint / cv::DMatch
construction (src/SuperPoint.cc:716, src/LightGlue.cc:352-361), slot ownership through shared_ptr deleters
(include/DescriptorPool.h:62-76), status -> empty result.  The same harness linked against the real library
(libdropin.so) is exercised without a GPU here (every initialize() fails loudly, every interface call returns empty,
nothing crashes) and on the GPU in tests/test_gpu_zz_dropin.py."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import frontend as ofe

FAKE = os.path.join(ROOT, "oracle", "_ref", "libdropin_fake.so")
REAL = os.path.join(ROOT, "oracle", "_ref", "libdropin.so")
CAP = 1024
fp, ip, dp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_double)


def bind(path):
    lib = C.CDLL(path)
    lib.dropin_create.restype = C.c_void_p
    lib.dropin_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_float, ip]
    lib.dropin_destroy.argtypes = [C.c_void_p]
    lib.dropin_process.restype = C.c_int
    lib.dropin_process.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                   C.c_int, fp, fp, fp, dp, C.c_char_p, ip]
    lib.dropin_promote_keyframe.restype = C.c_int
    lib.dropin_promote_keyframe.argtypes = [C.c_void_p, fp, C.c_int]
    for name in ("dropin_track", "dropin_verify"):
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_int, ip, ip, fp]
    lib.dropin_release_frames.argtypes = [C.c_void_p]
    lib.dropin_infer.restype = C.c_int
    lib.dropin_infer.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, fp, fp, fp]
    lib.dropin_hold_frame.restype = C.c_int
    lib.dropin_hold_frame.argtypes = [C.c_void_p]
    return lib


class Harness:
    """ctypes face of oracle/dropin_harness.cpp (shared with the GPU test)."""

    def __init__(self, lib, sp_w, lg_w, w, h, max_kp=CAP, thr=0.005, rb=4, min_disp=1.0):
        self.lib, self.cap = lib, max_kp
        st = C.c_int(-1)
        self.h = lib.dropin_create(sp_w.encode(), lg_w.encode(), w, h, max_kp, thr, rb, min_disp, C.byref(st))
        self.status = st.value

    def close(self):
        if self.h:
            self.lib.dropin_destroy(self.h)
            self.h = None

    def process(self, left, right, ts=0.0):
        left, right = np.asarray(left), np.asarray(right)
        assert left.dtype == np.uint8 and left.shape[:2] == right.shape[:2]
        ch = 1 if left.ndim == 2 else left.shape[2]
        ch_r = 1 if right.ndim == 2 else right.shape[2]
        cap = self.cap
        xy, resp, sa = np.zeros((cap, 2), np.float32), np.zeros(cap, np.float32), np.zeros((cap, 2), np.float32)
        stereo, has, desc = np.zeros((cap, 3), np.float64), np.zeros(cap, np.int8), np.zeros(4, np.int32)
        n = self.lib.dropin_process(self.h, left.ctypes.data, right.ctypes.data, left.shape[0], left.shape[1], left.strides[0],
                                    right.strides[0], ch + 16 * (ch_r if ch_r != ch else 0), ts, cap, xy.ctypes.data_as(fp), resp.ctypes.data_as(fp), sa.ctypes.data_as(fp),
                                    stereo.ctypes.data_as(dp), has.ctypes.data_as(C.c_char_p), desc.ctypes.data_as(ip))
        return dict(n=n, xy=xy[:n], response=resp[:n], size_angle=sa[:n], stereo=stereo[:n], has_depth=has[:n],
                    desc=dict(count=int(desc[0]), dim=int(desc[1]), slot=int(desc[2]), resident=bool(desc[3])))

    def infer(self, image):
        image = np.asarray(image)
        ch = 1 if image.ndim == 2 else image.shape[2]
        xy, resp, desc = np.zeros((self.cap, 2), np.float32), np.zeros(self.cap, np.float32), np.zeros((self.cap, 256), np.float32)
        n = self.lib.dropin_infer(self.h, image.ctypes.data, image.shape[0], image.shape[1], image.strides[0], ch, self.cap,
                                  xy.ctypes.data_as(fp), resp.ctypes.data_as(fp), desc.ctypes.data_as(fp))
        return n, xy[:max(n, 0)], resp[:max(n, 0)], desc[:max(n, 0)]

    def promote_keyframe(self):
        out = np.zeros((self.cap, 256), np.float32)
        rows = self.lib.dropin_promote_keyframe(self.h, out.ctypes.data_as(fp), self.cap)
        return rows, out[:max(rows, 0)]

    def _matches(self, fn):
        q, t, d = np.zeros(self.cap, np.int32), np.zeros(self.cap, np.int32), np.zeros(self.cap, np.float32)
        n = fn(self.h, self.cap, q.ctypes.data_as(ip), t.ctypes.data_as(ip), d.ctypes.data_as(fp))
        return q[:n], t[:n], d[:n]

    def track(self):
        return self._matches(self.lib.dropin_track)

    def verify(self):
        return self._matches(self.lib.dropin_verify)

    def release_frames(self):
        self.lib.dropin_release_frames(self.h)


# ---- the canned rules of oracle/fake_capi.cpp ----------------------------------------------------------------------
def fake_features(img, max_kp):
    flat = img.reshape(img.shape[0], -1)
    n = min(max_kp, 4 * int(flat[0, 0]))
    k = np.arange(n)
    xy = np.stack([flat[0, 1] + 2.0 * k, flat[1, 0] + (k % 7)], 1).astype(np.float32)
    return xy, (np.float32(1) / (np.float32(1) + k.astype(np.float32))).astype(np.float32), float(flat[0, 2])


def fake_matches(n0, n1):
    i = np.arange(n0)
    hit = (i % 3 != 0) & (n1 > 0)
    m = np.where(hit, (7 * i + 3) % max(n1, 1), -1).astype(np.int32)
    s = np.where(hit, 0.25 + 0.5 * (i % 2), 0.0).astype(np.float32)
    return m, s


def image(n_quarter, x0, y0, tag, h=6, w=40, pad=0, channels=1):
    """A tiny image whose first pixels drive the fake extractor; `pad` extra bytes per row test the stride."""
    buf = np.zeros((h, w * channels + pad), np.uint8)
    buf[0, 0], buf[0, 1], buf[0, 2], buf[1, 0] = n_quarter, x0, tag, y0
    view = buf[:, :w * channels]
    return view if channels == 1 else view.reshape(h, w, channels)


@pytest.fixture(scope="module")
def fake():
    if not os.path.exists(FAKE):
        pytest.skip("oracle/_ref/libdropin_fake.so not built (build() with /root/reference mounted)")
    lib = bind(FAKE)
    for name in ("fake_slots_in_use", "fake_release_calls", "fake_sp_destroyed", "fake_lg_alive"):
        getattr(lib, name).restype = C.c_int
    lib.fake_last_match.argtypes = [ip, dp]
    return lib


def last_match(lib):
    info, sums = np.zeros(6, np.int32), np.zeros(4, np.float64)
    lib.fake_last_match(info.ctypes.data_as(ip), sums.ctypes.data_as(dp))
    return dict(clone=info[0], device_path=info[1], n0=info[2], n1=info[3], slot0=info[4], slot1=info[5],
                xy0_sum=sums[0], xy1_sum=sums[1], d0_first=sums[2], d1_first=sums[3])


@pytest.mark.parametrize("pad,channels", [(0, 1), (24, 1), (8, 3)])
def test_reference_stereo_frontend_over_the_adapter(fake, pad, channels):
    hs = Harness(fake, "sp.ssbw", "lg.ssbw", 640, 480)
    assert hs.status == 7
    # left keypoints start at x = 60, right ones at x = 40, both step 2 px per index: the double's (7 i + 3) % n1
    # pairing yields positive, negative and sub-threshold disparities and row offsets -5 .. 7
    left, right = image(50, 60, 10, 3, pad=pad, channels=channels), image(45, 40, 9, 5, pad=pad, channels=channels)
    out = hs.process(left, right, 1.5)
    xl, sl, _ = fake_features(left, CAP)
    xr, _, _ = fake_features(right, CAP)
    assert out["n"] == 200 and np.array_equal(out["xy"], xl) and np.array_equal(out["response"], sl)
    assert np.all(out["size_angle"] == np.float32([1.0, -1.0]))            # src/SuperPoint.cc:716
    m0, s0 = fake_matches(len(xl), len(xr))
    q, t, _ = ofe.dmatches(m0, s0)
    stereo, has = ofe.stereo_postfilter(xl, xr, q, t)                       # restated src/StereoFrontEnd.cc:22-47
    assert np.array_equal(out["has_depth"], has) and 0 < has.sum() < len(q)
    assert np.array_equal(out["stereo"], stereo, equal_nan=True)
    lm = last_match(fake)
    assert (lm["clone"], lm["device_path"], lm["n0"], lm["n1"]) == (0, 1, 200, 180)
    assert lm["xy0_sum"] == float(xl.astype(np.float64).sum()) and lm["xy1_sum"] == float(xr.astype(np.float64).sum())
    assert lm["slot0"] == 0 and lm["slot1"] == 1                            # LIFO free list: L got slot 0, R slot 1
    # the frame keeps L's slot; R's handle died when process() returned (its shared_ptr deleter released slot 1)
    assert out["desc"] == dict(count=200, dim=256, slot=0, resident=True)
    assert fake.fake_slots_in_use() == 1
    hs.release_frames()
    assert fake.fake_slots_in_use() == 0
    hs.close()
    assert fake.fake_lg_alive() == 0


def test_left_and_right_images_with_different_row_strides(fake):
    """cv::Mat steps may differ between the two images (an ROI on one side): the C-ABI takes one stride per call, so
    the adapter packs them first - same features as with equal strides."""
    hs = Harness(fake, "sp.ssbw", "lg.ssbw", 640, 480)
    a = hs.process(image(20, 60, 10, 3, pad=0), image(20, 40, 9, 5, pad=0))
    b = hs.process(image(20, 60, 10, 3, pad=16), image(20, 40, 9, 5, pad=40))
    assert b["n"] == 80 and np.array_equal(a["xy"], b["xy"]) and np.array_equal(a["stereo"], b["stereo"], equal_nan=True)
    assert np.array_equal(a["has_depth"], b["has_depth"]) and a["has_depth"].sum() > 0
    hs.close()


def test_bgr_left_and_gray_right(fake):
    """The reference converts each side to gray on its own (src/SuperPoint.cc:768-773), so a colour / gray pair is legal."""
    hs = Harness(fake, "sp.ssbw", "lg.ssbw", 640, 480)
    out = hs.process(image(20, 60, 10, 3, channels=3), image(20, 40, 9, 5))
    assert out["n"] == 80 and out["has_depth"].sum() > 0 and fake.fake_slots_in_use() == 1
    hs.close()


def test_tracking_match_keyframe_record_and_loop_verification(fake):
    hs = Harness(fake, "sp.ssbw", "lg.ssbw", 640, 480)
    a = hs.process(image(30, 50, 8, 11), image(30, 38, 8, 12))
    rows, kf_desc = hs.promote_keyframe()                                   # last_keyframe_ = frame + descriptors_to_host
    assert rows == 120 and kf_desc.shape == (120, 256)
    k, c = np.arange(120, dtype=np.float32)[:, None], np.arange(256, dtype=np.float32)[None, :]
    assert np.array_equal(kf_desc, (np.float32(11) + k + c / np.float32(1024)).astype(np.float32))
    assert fake.fake_slots_in_use() == 1                                    # frame and last_keyframe_ share slot 0
    b = hs.process(image(25, 52, 8, 21), image(25, 40, 8, 22))              # next frame: L takes the free slot 1
    assert b["desc"]["slot"] == 1 and fake.fake_slots_in_use() == 2         # keyframe's slot 0 stayed alive
    q, t, d = hs.track()                                                    # src/VoEstimator.cc:240-246
    lm = last_match(fake)
    assert (lm["clone"], lm["device_path"], lm["n0"], lm["n1"], lm["slot0"], lm["slot1"]) == (0, 1, 120, 100, 0, 1)
    m0, s0 = fake_matches(120, 100)
    eq, et, ed = ofe.dmatches(m0, s0)
    assert np.array_equal(q, eq) and np.array_equal(t, et) and np.array_equal(d, ed)   # distance = 1 - score
    assert np.all(np.diff(q) > 0)                                           # increasing queryIdx
    q2, t2, d2 = hs.verify()                                                # src/LoopCloser.cc:44-53 on the loop matcher
    lm = last_match(fake)
    assert (lm["clone"], lm["device_path"], lm["n0"], lm["n1"]) == (1, 0, 120, 100)
    assert lm["d0_first"] == 11.0 and lm["d1_first"] == 21.0 and lm["xy0_sum"] == float(a["xy"].astype(np.float64).sum())
    assert np.array_equal(q2, eq) and np.array_equal(t2, et) and np.array_equal(d2, ed)
    releases = fake.fake_release_calls()
    hs.release_frames()
    assert fake.fake_slots_in_use() == 0 and fake.fake_release_calls() == releases + 2   # one release per slot, not per copy
    hs.close()


def test_pool_exhaustion_failed_inference_and_empty_images(fake):
    hs = Harness(fake, "sp.ssbw", "lg.ssbw", 640, 480, max_kp=64)
    l, r = image(10, 30, 5, 1), image(10, 20, 5, 2)
    for i in range(7):                                                      # seven frames kept alive: slots 0..6 stay out
        out = hs.process(l, r)
        assert out["desc"]["slot"] == i and out["has_depth"].sum() > 0
        assert hs.lib.dropin_hold_frame(hs.h) == i + 1
    assert fake.fake_slots_in_use() == 7
    # eighth frame: L takes the last slot, R finds the pool exhausted -> keypoints without descriptors
    # (src/SuperPoint.cc:724-727), so the device match sees an empty handle and returns nothing: no stereo depth
    out = hs.process(l, r)
    assert out["n"] == 40 and out["desc"]["slot"] == 7 and out["has_depth"].sum() == 0
    hs.lib.dropin_hold_frame(hs.h)
    # ninth: L itself gets no slot - the handle DescriptorPool::make returns when exhausted (count set, slot -1, no data)
    out = hs.process(l, r)
    assert out["n"] == 40 and out["desc"] == dict(count=40, dim=256, slot=-1, resident=False)
    assert fake.fake_slots_in_use() == 8 and hs.promote_keyframe()[0] == 0  # empty handle -> empty Mat
    hs.release_frames()
    assert fake.fake_slots_in_use() == 0
    out = hs.process(l, r)
    assert out["desc"]["slot"] == 7 and out["has_depth"].sum() > 0          # LIFO: the slot released last comes back first
    hs.promote_keyframe()
    # a failed inference (status != OK) -> empty Features -> empty frame, no crash (src/SuperPoint.cc:895-899)
    bad = hs.process(image(255, 30, 5, 1), r)
    assert bad["n"] == 0 and bad["desc"] == dict(count=0, dim=0, slot=-1, resident=False)
    assert len(hs.track()[0]) == 0                                          # keyframe <-> empty frame: no matches
    # zero keypoints on one side -> the matcher returns an empty result -> no depth anywhere
    none_right = hs.process(l, image(0, 20, 5, 2))
    assert none_right["n"] == 40 and none_right["has_depth"].sum() == 0 and np.isnan(none_right["stereo"][:, 1]).all()
    hs.close()
    assert fake.fake_slots_in_use() == 0


def test_host_path_infer_of_the_demo_programs(fake):
    """SuperPoint::infer (tests/test_superpoint_only.cc:71): keypoints + CV_32F descriptors on the host, slot returned."""
    hs = Harness(fake, "sp.ssbw", "lg.ssbw", 640, 480, max_kp=64)
    img = image(9, 17, 3, 40, pad=5)
    n, xy, resp, desc = hs.infer(img)
    exy, esc, tag = fake_features(img, 64)
    assert n == 36 and np.array_equal(xy, exy) and np.array_equal(resp, esc)
    k, c = np.arange(36, dtype=np.float32)[:, None], np.arange(256, dtype=np.float32)[None, :]
    assert np.array_equal(desc, (np.float32(tag) + k + c / np.float32(1024)).astype(np.float32))
    assert fake.fake_slots_in_use() == 0                                    # the transient handle gave its slot back
    assert hs.infer(image(0, 17, 3, 40))[0] == 0                            # no keypoints: success, empty outputs
    assert hs.infer(image(255, 17, 3, 40))[0] == -1                         # failed inference -> false
    hs.close()
    broken = Harness(fake, "missing", "lg.ssbw", 640, 480)
    assert broken.infer(img)[0] == -1
    broken.close()


def test_initialize_failure_leaves_a_broken_but_safe_object(fake):
    hs = Harness(fake, "missing", "missing", 640, 480)
    assert hs.status == 0
    out = hs.process(image(10, 30, 5, 1), image(10, 20, 5, 2))
    assert out["n"] == 0 and hs.promote_keyframe()[0] == 0 and len(hs.track()[0]) == 0 and len(hs.verify()[0]) == 0
    hs.close()


@pytest.mark.skipif(not os.path.exists(REAL), reason="oracle/_ref/libdropin.so not built")
def test_real_library_without_a_gpu_fails_loudly_and_returns_empty():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: the real path is covered by tests/test_gpu_zz_dropin.py")
    hs = Harness(bind(REAL), "w.ssbw", "w.ssbw", 640, 480)
    assert hs.status == 0                                                   # no device -> every initialize() is false
    img = np.zeros((480, 640), np.uint8)
    out = hs.process(img, img)
    assert out["n"] == 0 and hs.promote_keyframe()[0] == 0 and len(hs.track()[0]) == 0
    hs.close()


# ---- the other adapter classes (EigenPlacesB200, RemapB200, RgbdPostB200) over the C-ABI double -----------------------
u8p = C.c_void_p


def test_place_recognizer_adapter_marshalling_and_index_semantics(fake, monkeypatch):
    from oracle import eigenplaces as oep

    fake.dropin_place_create.restype = C.c_void_p
    fake.dropin_place_create.argtypes = [C.c_char_p, C.c_int, C.c_int, ip]
    fake.dropin_place_destroy.argtypes = [C.c_void_p]
    fake.dropin_place_compute.restype = C.c_int
    fake.dropin_place_compute.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, C.c_int, C.c_int, fp]
    fake.dropin_place_add.argtypes = [C.c_void_p, C.c_size_t, fp, C.c_int, C.c_int]
    fake.dropin_place_query.restype = C.c_int
    fake.dropin_place_query.argtypes = [C.c_void_p, fp, C.c_int, C.c_size_t, C.c_int, C.POINTER(C.c_size_t), fp]
    ok = C.c_int(-1)
    bad = fake.dropin_place_create(b"missing", 512, 512, C.byref(ok))
    img = np.arange(6 * 50 * 3, dtype=np.uint8).reshape(6, 50, 3)
    out = np.zeros(512, np.float32)
    assert ok.value == 0 and fake.dropin_place_compute(bad, img.ctypes.data, 6, 50, img.strides[0], 3, out.ctypes.data_as(fp)) == 0
    fake.dropin_place_destroy(bad)
    monkeypatch.setenv("SUPERSLAM_LOOP_MIN_SCORE", "0.30")                  # src/EigenPlaces.cc:30-34
    ep = fake.dropin_place_create(b"ep.ssbw", 512, 512, C.byref(ok))
    assert ok.value == 1
    # image -> (pointer, rows, cols, step, channels): a padded BGR view; the double echoes row 1
    buf = np.random.default_rng(0).integers(0, 250, (6, 50 * 3 + 7), dtype=np.uint8)
    view = buf[:, :150].reshape(6, 50, 3)
    assert fake.dropin_place_compute(ep, view.ctypes.data, 6, 50, buf.strides[0], 3, out.ctypes.data_as(fp)) == 512
    assert np.array_equal(out, buf[1, np.arange(512) % 150].astype(np.float32) + 1)
    blank = np.full((6, 50), 255, np.uint8)
    assert fake.dropin_place_compute(ep, blank.ctypes.data, 6, 50, 50, 1, out.ctypes.data_as(fp)) == 0   # failure -> empty Mat
    # add / query against the restated CosineDescriptorIndex (src/PlaceRecognizer.cc:21-52)
    rng = np.random.default_rng(1)
    base = rng.normal(size=512).astype(np.float32)
    index = oep.CosineDescriptorIndex()
    for k in range(40):
        d = (base + rng.normal(size=512) * (0.2 + 0.1 * k)).astype(np.float32)
        fake.dropin_place_add(ep, 100 + k, d.ctypes.data_as(fp), 512, k % 2)            # every other one as a CV_64F column
        index.add(100 + k, d)
    ids, sc = (C.c_size_t * 64)(), np.zeros(64, np.float32)
    for exclude, top_k in [(0, 5), (10, 3), (39, 0), (40, 5), (5, 0)]:
        n = fake.dropin_place_query(ep, base.ctypes.data_as(fp), 512, exclude, top_k, ids, sc.ctypes.data_as(fp))
        exp = index.query(base, exclude, top_k, 0.30)
        assert n == len(exp) and [ids[i] for i in range(n)] == [e[0] for e in exp]
        assert np.allclose(sc[:n], [e[1] for e in exp], atol=1e-6)
    assert len(index.query(base, 5, 0, 0.30)) < 35                          # the environment threshold really filters
    fake.dropin_place_destroy(ep)


def test_remap_adapter_creates_the_destination_and_passes_the_stride(fake):
    fake.dropin_remap.restype = C.c_int
    fake.dropin_remap.argtypes = [fp, fp, C.c_int, C.c_int, u8p, C.c_int, C.c_int, C.c_int, u8p]
    fake.fake_rect_stride.restype = C.c_int
    dh, dw, sh, sw = 10, 12, 7, 9
    mx, my = np.zeros((dh, dw), np.float32), np.zeros((dh, dw), np.float32)
    src = np.random.default_rng(2).integers(0, 200, (sh, sw + 5), dtype=np.uint8)
    dst = np.zeros((dh, dw), np.uint8)
    assert fake.dropin_remap(mx.ctypes.data_as(fp), my.ctypes.data_as(fp), dh, dw, src.ctypes.data, sh, sw, src.strides[0],
                             dst.ctypes.data) == 1
    y, x = np.mgrid[0:dh, 0:dw]
    assert fake.fake_rect_stride() == sw + 5 and np.array_equal(dst, src[y % sh, x % sw] + 1)
    # a map whose size the library rejects (not a multiple of four pixels): the functor reports failure
    assert fake.dropin_remap(mx.ctypes.data_as(fp), my.ctypes.data_as(fp), 3, 3, src.ctypes.data, sh, sw, src.strides[0],
                             dst.ctypes.data) == 0


def test_rgbd_post_without_a_distortion_model_and_place_recognizer_with_empty_descriptors(fake):
    """Empty cv::Mat arguments: cv::Mat::reshape throws on them in OpenCV (the stand-in aborts), so the adapter must
    not reshape an empty dist_coeffs (RgbdFrontEnd allows none, src/RgbdFrontEnd.cc:30) or an empty descriptor."""
    fake.dropin_rgbd_post.restype = C.c_int
    fake.dropin_rgbd_post.argtypes = [fp, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, C.c_int, C.c_int,
                                      C.c_double, C.c_double, C.c_double, fp, dp, C.c_char_p]
    fake.fake_last_rgbd.argtypes = [dp]
    xy = np.float32([[3, 4], [5.25, 6.5]])
    depth = np.full((10, 12), 2000, np.uint16)
    cam = np.array([500.0, 500.0, 6.0, 5.0])
    oxy, ost, ohas = np.zeros((2, 2), np.float32), np.zeros((2, 3)), np.zeros(2, np.int8)
    assert fake.dropin_rgbd_post(xy.ctypes.data_as(fp), 2, depth.ctypes.data, 0, 10, 12, 24, cam.ctypes.data_as(dp), None, 0, 0,
                                 40.0, 5000.0, 8.0, oxy.ctypes.data_as(fp), ost.ctypes.data_as(dp),
                                 ohas.ctypes.data_as(C.c_char_p)) == 1
    g = np.zeros(28)
    fake.fake_last_rgbd(g.ctypes.data_as(dp))
    assert g[9] == 0 and g[13] == 0 and np.all(g[14:28] == -1) and ohas.tolist() == [1, 1]     # n_dist 0, dist NULL
    # no depth image at all: every sample is 0 (sampleDepth's bounds check) - a 1 x 1 zero map says the same
    assert fake.dropin_rgbd_post(xy.ctypes.data_as(fp), 2, None, 1, 0, 0, 0, cam.ctypes.data_as(dp), None, 0, 0,
                                 40.0, 5000.0, 8.0, oxy.ctypes.data_as(fp), ost.ctypes.data_as(dp),
                                 ohas.ctypes.data_as(C.c_char_p)) == 1
    fake.fake_last_rgbd(g.ctypes.data_as(dp))
    assert list(g[1:5]) == [0, 1, 1, 2] and ohas.tolist() == [0, 0]
    fake.dropin_place_create.restype = C.c_void_p
    fake.dropin_place_create.argtypes = [C.c_char_p, C.c_int, C.c_int, ip]
    fake.dropin_place_add.argtypes = [C.c_void_p, C.c_size_t, fp, C.c_int, C.c_int]
    fake.dropin_place_query.restype = C.c_int
    fake.dropin_place_query.argtypes = [C.c_void_p, fp, C.c_int, C.c_size_t, C.c_int, C.POINTER(C.c_size_t), fp]
    ok = C.c_int()
    ep = fake.dropin_place_create(b"ep.ssbw", 512, 512, C.byref(ok))
    ids, sc = (C.c_size_t * 4)(), np.zeros(4, np.float32)
    fake.dropin_place_add(ep, 7, None, 0, 0)                                # add(id, cv::Mat()) - ignored
    assert fake.dropin_place_query(ep, None, 0, 0, 3, ids, sc.ctypes.data_as(fp)) == 0
    fake.dropin_place_destroy.argtypes = [C.c_void_p]
    fake.dropin_place_destroy(ep)


@pytest.mark.parametrize("depth_type,dist_f32", [(0, 0), (1, 1), (2, 0)])
def test_rgbd_post_adapter_marshalling(fake, depth_type, dist_f32):
    fake.dropin_rgbd_post.restype = C.c_int
    fake.dropin_rgbd_post.argtypes = [fp, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, C.c_int, C.c_int,
                                      C.c_double, C.c_double, C.c_double, fp, dp, C.c_char_p]
    fake.fake_last_rgbd.argtypes = [dp]
    rng = np.random.default_rng(3)
    n, dh, dw = 50, 20, 30
    xy = (rng.uniform(0, 1, (n, 2)) * [dw + 2, dh + 2] - 1).astype(np.float32)
    dtype = [np.uint16, np.float32, np.uint8][depth_type]
    dbuf = rng.integers(0, 4, (dh, dw + 3)).astype(dtype) * (1000 if depth_type != 2 else 1)
    depth = dbuf[:, :dw]
    cam = np.array([500.5, 499.25, 15.5, 9.75])
    dist = np.array([-0.25, 0.0625, 0.001953125, -0.00048828125, 0.0078125])       # exactly representable in fp32
    oxy, ost, ohas = np.zeros((n, 2), np.float32), np.zeros((n, 3)), np.zeros(n, np.int8)
    assert fake.dropin_rgbd_post(xy.ctypes.data_as(fp), n, dbuf.ctypes.data, depth_type, dh, dw, dbuf.strides[0],
                                 cam.ctypes.data_as(dp), dist.ctypes.data_as(dp), 5, dist_f32, 40.0, 5000.0, 8.0,
                                 oxy.ctypes.data_as(fp), ost.ctypes.data_as(dp), ohas.ctypes.data_as(C.c_char_p)) == 1
    g = np.zeros(28)
    fake.fake_last_rgbd(g.ctypes.data_as(dp))
    # an unsupported depth type is replaced by an all-zero CV_16U map (sampleDepth returns 0, src/RgbdFrontEnd.cc:12-20)
    exp_type, exp_stride = (depth_type, dbuf.strides[0]) if depth_type != 2 else (0, dw * 2)
    assert list(g[:5]) == [n, exp_type, dh, dw, exp_stride] and np.array_equal(g[5:9], cam)
    assert list(g[9:14]) == [5, 40.0, 5000.0, 8.0, 1] and np.array_equal(g[14:19], dist) and np.all(g[19:28] == -1)
    u, v = np.rint(xy[:, 0]).astype(int), np.rint(xy[:, 1]).astype(int)     # no .5 ties in this sample
    inside = (u >= 0) & (v >= 0) & (u < dw) & (v < dh)
    z = np.where(inside, depth[np.clip(v, 0, dh - 1), np.clip(u, 0, dw - 1)], 0).astype(np.float64)
    if depth_type == 2:
        z[:] = 0
    assert np.array_equal(oxy, xy + np.float32(0.25)) and np.array_equal(ohas, (z != 0).astype(np.int8))
    assert np.array_equal(ost, np.stack([xy[:, 0].astype(np.float64), z / 5000.0, xy[:, 1].astype(np.float64)], 1))
