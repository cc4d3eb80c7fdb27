"""The REFERENCE's whole wrapper path for one stereo frame, run end to end on the CPU: src/SuperPoint.cc, src/LightGlue.cc,
src/DescriptorPool.cc and src/StereoFrontEnd.cc compiled in place, unchanged (oracle/ref_e2e_shim.cpp ->
oracle/_ref/libref_e2e.so), over a functional TensorRT stand-in whose enqueueV3 hands the bound buffers to this test -
which serves the two networks with the CPU oracle -, the CUDA runtime calls on host memory and the gather kernel as a
callback.  Everything else between the images and the StereoFrame is the reference's own code: gray / 255 conversion,
{2,1,H,W} packing, buffer (re)sizing, slice offsets, select_and_gather, pool slots, keypoint normalisation, fp16 descriptor
copies, postprocess_outputs, the disparity / row filter.  The oracle's composition of the restated pieces - the thing
every GPU parity test compares the product with - must reproduce it bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from oracle import frontend as ofe
from oracle import lightglue as olg
from oracle import superpoint as osp

LIB = os.path.join(ROOT, "oracle", "_ref", "libref_e2e.so")
pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libref_e2e.so not built")
fp, ip, u16p, u8p = C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_uint16), C.POINTER(C.c_uint8)
SP_FN = C.CFUNCTYPE(None, fp, C.c_int, C.c_int, C.c_int, fp, u16p)
LG_FN = C.CFUNCTYPE(None, fp, C.c_int, u16p, fp, C.c_int, u16p, ip, fp)
GATHER_FN = C.CFUNCTYPE(None, u16p, C.c_int, C.c_int, C.c_int, ip, ip, C.c_int, u16p)


def arr(ptr, shape, dtype):
    return np.ctypeslib.as_array(ptr, shape=(int(np.prod(shape)),)).view(dtype).reshape(shape)


@pytest.fixture(scope="module")
def ref(sp_weights, lg_weights, tmp_path_factory):
    lib = C.CDLL(LIB)
    lib.ref_e2e_create.restype = C.c_void_p
    lib.ref_e2e_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_float, ip]
    lib.ref_e2e_destroy.argtypes = [C.c_void_p]
    lib.ref_e2e_live_allocations.restype = C.c_long
    lib.ref_e2e_process.restype = C.c_int
    lib.ref_e2e_process.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, ip, fp, fp, u16p, fp,
                                    fp, u16p, C.POINTER(C.c_double), C.c_char_p, ip, ip, fp]
    calls = {"sp": [], "lg": [], "gather": 0, "module": None}

    @SP_FN
    def sp_infer(image, b, h, w, scores, desc):
        x = arr(image, (b, 1, h, w), np.float32).copy()
        if calls["module"] is not None:   # the reference's own torch module as the engine (the graph the ONNX is traced from)
            import torch

            with torch.no_grad():
                s, grid = (t.numpy() for t in calls["module"](torch.from_numpy(x)))
        else:
            s, grid, _ = osp.dense_forward(x, sp_weights, fp16_storage=False)   # the restated graph
        arr(scores, s.shape, np.float32)[:] = s
        arr(desc, grid.shape, np.uint16)[:] = grid.astype(np.float16).view(np.uint16)   # the engine's fp16 binding
        calls["sp"].append((b, h, w))

    @LG_FN
    def lg_infer(k0, n0, d0, k1, n1, d1, m0, ms0):
        a = olg.match(lg_weights, arr(k0, (n0, 2), np.float32).copy(), arr(d0, (n0, 256), np.uint16).view(np.float16).copy(),
                      arr(k1, (n1, 2), np.float32).copy(), arr(d1, (n1, 256), np.uint16).view(np.float16).copy())
        arr(m0, (n0,), np.int32)[:] = a[0]
        arr(ms0, (n0,), np.float32)[:] = a[1]
        calls["lg"].append((n0, n1))

    @GATHER_FN
    def gather(grid, c, gh, gw, cell_h, cell_w, n, out):
        cell = np.stack([arr(cell_h, (n,), np.int32), arr(cell_w, (n,), np.int32)], 1)
        rows = osp.gather_normalize(arr(grid, (c, gh, gw), np.uint16).view(np.float16).copy(), cell)
        arr(out, (n, c), np.uint16)[:] = rows.view(np.uint16)
        calls["gather"] += 1

    lib.ref_e2e_set_hooks(sp_infer, lg_infer, gather)
    d = tmp_path_factory.mktemp("engines")
    (d / "sp.engine").write_bytes(b"superpoint stand-in engine")
    (d / "lg.engine").write_bytes(b"lightglue stand-in engine")
    (d / "other.engine").write_bytes(b"built by another TensorRT")
    return dict(lib=lib, calls=calls, dir=d, keep=(sp_infer, lg_infer, gather))


def run(ref, left, right, K, thr=0.005, rb=4, min_disp=1.0):
    lib = ref["lib"]
    h, w = left.shape
    st = C.c_int(-1)
    hnd = lib.ref_e2e_create(str(ref["dir"] / "sp.engine").encode(), str(ref["dir"] / "lg.engine").encode(), K, thr, rb, w, h,
                             min_disp, C.byref(st))
    assert st.value == 3
    cap = K
    counts = np.zeros(3, np.int32)
    xl, rl, dl = np.zeros((cap, 2), np.float32), np.zeros(cap, np.float32), np.zeros((cap, 256), np.uint16)
    xr, rr, dr = np.zeros((cap, 2), np.float32), np.zeros(cap, np.float32), np.zeros((cap, 256), np.uint16)
    stereo, has = np.zeros((cap, 3)), np.zeros(cap, np.int8)
    q, t, dist = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.float32)
    n = lib.ref_e2e_process(hnd, left.ctypes.data, right.ctypes.data, h, w, left.strides[0], cap, counts.ctypes.data_as(ip),
                            xl.ctypes.data_as(fp), rl.ctypes.data_as(fp), dl.ctypes.data_as(u16p), xr.ctypes.data_as(fp),
                            rr.ctypes.data_as(fp), dr.ctypes.data_as(u16p), stereo.ctypes.data_as(C.POINTER(C.c_double)),
                            has.ctypes.data_as(C.c_char_p), q.ctypes.data_as(ip), t.ctypes.data_as(ip), dist.ctypes.data_as(fp))
    lib.ref_e2e_destroy(hnd)
    nl, nr, nm = (int(v) for v in counts)
    assert n == nl
    return dict(xy=(xl[:nl], xr[:nr]), score=(rl[:nl], rr[:nr]), desc=(dl[:nl].view(np.float16), dr[:nr].view(np.float16)),
                stereo=stereo[:nl], has_depth=has[:nl], query=q[:nm], train=t[:nm], distance=dist[:nm])


def oracle_frame(left, right, sp_weights, lg_weights, K, thr=0.005, rb=4, min_disp=1.0):
    h, w = left.shape
    f = osp.extract(np.stack([left, right]), sp_weights, K, thr, rb)
    m0, ms0 = olg.match(lg_weights, olg.normalize_keypoints(f[0]["xy"], w, h), f[0]["desc"],
                        olg.normalize_keypoints(f[1]["xy"], w, h), f[1]["desc"])
    q, t, d = ofe.dmatches(m0, ms0)
    stereo, has = ofe.stereo_postfilter(f[0]["xy"], f[1]["xy"], q, t, min_disp)
    return f, q, t, d, stereo, has


@pytest.mark.parametrize("case", ["golden_pair", "odd_size_shifted", "few_keypoints"])
def test_reference_wrappers_end_to_end_equal_the_oracle_composition(ref, sp_weights, lg_weights, case):
    from superslam_b200.synth import synth_pair

    if case == "golden_pair":           # the images behind tests/golden/superpoint_ref_small.npz
        imgs = np.load(os.path.join(GOLDEN, "superpoint_ref_small.npz"))["images"]
        left, right, K = np.ascontiguousarray(imgs[0]), np.ascontiguousarray(imgs[1]), 256
    elif case == "odd_size_shifted":    # 99 x 131: floor-halved pools, 96 x 128 score map, scale factors != 1; padded rows
        l, r = synth_pair(99, 131, 21, 60)
        buf = np.zeros((2, 99, 131 + 29), np.uint8)
        buf[:, :, :131] = [l, r]
        left, right, K = buf[0, :, :131], buf[1, :, :131], 128
    else:                               # K far below the candidate count: the top-K cut and the tie order matter
        left, right = synth_pair(120, 160, 5, 50)
        K = 24
    ref["calls"]["sp"].clear(), ref["calls"]["lg"].clear()
    got = run(ref, left, right, K)
    f, q, t, d, stereo, has = oracle_frame(np.ascontiguousarray(left), np.ascontiguousarray(right), sp_weights, lg_weights, K)
    h, w = left.shape
    assert ref["calls"]["sp"] == [(2, h, w), (2, h, w)]                     # one batched {2,1,H,W} pass per extract_stereo
    for i in range(2):
        assert len(got["xy"][i]) == len(f[i]["xy"]) > 10
        assert np.array_equal(got["xy"][i], f[i]["xy"]) and np.array_equal(got["score"][i], f[i]["score"])
        assert np.array_equal(got["desc"][i].view(np.uint16), f[i]["desc"].view(np.uint16))
    assert ref["calls"]["lg"] == [(len(f[0]["xy"]), len(f[1]["xy"]))] * 2
    assert np.array_equal(got["query"], q) and np.array_equal(got["train"], t) and np.array_equal(got["distance"], d)
    assert np.array_equal(got["has_depth"], has) and np.array_equal(got["stereo"], stereo, equal_nan=True)
    if case != "few_keypoints":
        assert len(q) > 5
    assert ref["lib"].ref_e2e_live_allocations() == 0                       # every buffer and pool slot was returned


@pytest.mark.skipif(not os.path.exists("/root/reference/utils/convert_superpoint_to_onnx.py"),
                    reason="the reference's torch module is not on this machine")
def test_reference_wrappers_over_the_reference_torch_module(ref, sp_weights, lg_weights):
    """No restated piece on the SuperPoint side: the reference's C++ wrapper code around the reference's OWN torch graph
    (utils/convert_superpoint_to_onnx.py: SuperPoint + DenseSuperPoint(nms_radius 4), weights/superpoint_v1.pth - what the
    TensorRT engine is built from) - i.e. the reference itself, run here in fp32.  The oracle reproduces its frame bit for bit."""
    import sys

    import torch

    from superslam_b200.synth import synth_pair

    sys.path.insert(0, "/root/reference/utils")
    import convert_superpoint_to_onnx as refmod

    net = refmod.SuperPoint()
    net.load_state_dict(torch.load("/root/reference/weights/superpoint_v1.pth", map_location="cpu", weights_only=True))
    ref["calls"]["module"] = refmod.DenseSuperPoint(net.eval(), 4).eval()
    try:
        left, right = synth_pair(240, 320, 1234, 120)
        K = 512
        got = run(ref, left, right, K)
    finally:
        ref["calls"]["module"] = None
    f, q, t, d, stereo, has = oracle_frame(left, right, sp_weights, lg_weights, K)
    for i in range(2):
        assert len(got["xy"][i]) == len(f[i]["xy"]) > 100
        assert np.array_equal(got["xy"][i], f[i]["xy"]) and np.array_equal(got["score"][i], f[i]["score"])
        assert np.array_equal(got["desc"][i].view(np.uint16), f[i]["desc"].view(np.uint16))
    assert np.array_equal(got["query"], q) and np.array_equal(got["train"], t) and np.array_equal(got["distance"], d)
    assert np.array_equal(got["has_depth"], has) and np.array_equal(got["stereo"], stereo, equal_nan=True) and has.sum() > 10


def test_mono_extract_descriptors_to_host_and_host_descriptor_match(ref, sp_weights, lg_weights):
    """The other call shapes of the interfaces, through the reference's own code: IFeatureExtractor::extract (batch-1
    dynamic shape, preprocess_image), IFeatureMatcher::descriptors_to_host, and the host-descriptor match of loop closure
    (CV_32F rows -> fp16 binding in prepare_inputs)."""
    from superslam_b200.synth import synth_pair

    lib = ref["lib"]
    lib.ref_e2e_mono_and_host_match.restype = C.c_int
    lib.ref_e2e_mono_and_host_match.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, ip, fp, fp,
                                                fp, fp, fp, fp, ip, ip, fp]
    a, b = synth_pair(120, 160, 9, 50)
    b = np.ascontiguousarray(np.roll(a, 3, axis=0))                         # the "same place" seen again, shifted
    K, h, w = 200, 120, 160
    st = C.c_int(-1)
    hnd = lib.ref_e2e_create(str(ref["dir"] / "sp.engine").encode(), str(ref["dir"] / "lg.engine").encode(), K, 0.005, 4, w, h,
                             1.0, C.byref(st))
    assert st.value == 3
    counts = np.zeros(3, np.int32)
    xy, resp, dh = [np.zeros((K, 2), np.float32) for _ in range(2)], [np.zeros(K, np.float32) for _ in range(2)], \
        [np.zeros((K, 256), np.float32) for _ in range(2)]
    q, t, dist = np.zeros(K, np.int32), np.zeros(K, np.int32), np.zeros(K, np.float32)
    ref["calls"]["sp"].clear(), ref["calls"]["lg"].clear()
    n = lib.ref_e2e_mono_and_host_match(hnd, a.ctypes.data, b.ctypes.data, h, w, a.strides[0], K, counts.ctypes.data_as(ip),
                                        xy[0].ctypes.data_as(fp), resp[0].ctypes.data_as(fp), dh[0].ctypes.data_as(fp),
                                        xy[1].ctypes.data_as(fp), resp[1].ctypes.data_as(fp), dh[1].ctypes.data_as(fp),
                                        q.ctypes.data_as(ip), t.ctypes.data_as(ip), dist.ctypes.data_as(fp))
    lib.ref_e2e_destroy(hnd)
    n0, n1, nm = (int(v) for v in counts)
    assert n == n0 > 10 and ref["calls"]["sp"] == [(1, h, w), (1, h, w)]    # two batch-1 passes
    f = [osp.extract(img[None], sp_weights, K)[0] for img in (a, b)]
    for i, cnt in enumerate((n0, n1)):
        assert cnt == len(f[i]["xy"]) and np.array_equal(xy[i][:cnt], f[i]["xy"]) and np.array_equal(resp[i][:cnt], f[i]["score"])
        assert np.array_equal(dh[i][:cnt], f[i]["desc"].astype(np.float32))                # fp16 slot widened to CV_32F
    m0, ms0 = olg.match(lg_weights, olg.normalize_keypoints(f[0]["xy"], w, h), f[0]["desc"],
                        olg.normalize_keypoints(f[1]["xy"], w, h), f[1]["desc"])
    eq, et, ed = ofe.dmatches(m0, ms0)
    assert nm == len(eq) > 5 and np.array_equal(q[:nm], eq) and np.array_equal(t[:nm], et) and np.array_equal(dist[:nm], ed)
    assert lib.ref_e2e_live_allocations() == 0


def test_a_foreign_engine_file_fails_initialize_like_a_tensorrt_mismatch(ref):
    lib = ref["lib"]
    st = C.c_int(-1)
    h = lib.ref_e2e_create(str(ref["dir"] / "other.engine").encode(), str(ref["dir"] / "missing.engine").encode(), 64, 0.005, 4,
                           160, 120, 1.0, C.byref(st))
    assert st.value == 0
    lib.ref_e2e_destroy(h)


def test_adapter_classes_equal_the_reference_classes_over_the_same_networks(ref, sp_weights, lg_weights):
    """Differential check of the drop-in boundary: StereoFrontEnd::process over the REFERENCE's SuperPoint / LightGlue classes
    (fake TensorRT, networks = oracle) against StereoFrontEnd::process over the ADAPTER's SuperPointB200 / LightGlueB200
    (C-ABI double in served mode, the same oracle behind ssb_sp_extract / ssb_lg_match_device): the frames must be identical -
    keypoints, responses, stereo points, depth flags, the descriptor handle's shape and the rows it points to."""
    from superslam_b200.synth import synth_pair
    from test_dropin_adapter import FAKE, Harness, bind

    if not os.path.exists(FAKE):
        pytest.skip("oracle/_ref/libdropin_fake.so not built")
    left, right = synth_pair(120, 160, 33, 60)
    h, w, K = 120, 160, 160
    want = run(ref, left, right, K)

    fake = bind(FAKE)
    EX = C.CFUNCTYPE(C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.c_int, fp, fp, u16p)
    MT = C.CFUNCTYPE(None, fp, C.c_int, u16p, fp, C.c_int, u16p, C.c_int, C.c_int, ip, fp)
    pair = osp.extract(np.stack([left, right]), sp_weights, K)              # one batch-2 pass, like extract_stereo
    by_image = {left.tobytes(): pair[0], right.tobytes(): pair[1]}

    @EX
    def extract(image, hh, ww, stride, max_kp, xy, score, desc):
        img = np.ctypeslib.as_array(image, shape=(hh, stride))[:, :ww]
        f = by_image[np.ascontiguousarray(img).tobytes()]
        n = len(f["xy"])
        assert n <= max_kp
        arr(xy, (n, 2), np.float32)[:] = f["xy"]
        arr(score, (n,), np.float32)[:] = f["score"]
        arr(desc, (n, 256), np.uint16)[:] = f["desc"].view(np.uint16)
        return n

    @MT
    def match(k0, n0, d0, k1, n1, d1, iw, ih, m0, ms0):
        a = olg.match(lg_weights, olg.normalize_keypoints(arr(k0, (n0, 2), np.float32).copy(), iw, ih),
                      arr(d0, (n0, 256), np.uint16).view(np.float16).copy(),
                      olg.normalize_keypoints(arr(k1, (n1, 2), np.float32).copy(), iw, ih),
                      arr(d1, (n1, 256), np.uint16).view(np.float16).copy())
        arr(m0, (n0,), np.int32)[:] = a[0]
        arr(ms0, (n0,), np.float32)[:] = a[1]

    fake.fake_set_servers.argtypes = [EX, MT]
    fake.fake_set_servers(extract, match)
    try:
        hs = Harness(fake, "sp.ssbw", "lg.ssbw", w, h, max_kp=K)
        got = hs.process(left, right, 0.0)
        rows, host = hs.promote_keyframe()
        hs.release_frames()
        hs.close()
    finally:
        fake.fake_set_servers(C.cast(None, EX), C.cast(None, MT))
    n = len(want["xy"][0])
    assert got["n"] == n > 20 and np.array_equal(got["xy"], want["xy"][0]) and np.array_equal(got["response"], want["score"][0])
    assert np.array_equal(got["has_depth"], want["has_depth"]) and got["has_depth"].sum() > 3
    assert np.array_equal(got["stereo"], want["stereo"], equal_nan=True)
    assert got["desc"] == dict(count=n, dim=256, slot=0, resident=True)     # L took the first slot in both object graphs
    assert rows == n and np.array_equal(host, want["desc"][0].astype(np.float32))
