"""The REFERENCE's own CUDA kernel (src/DescriptorGather.cu, compiled in place by oracle/Makefile into
oracle/_ref/libref_gather.so) against the product's fused gather, on the product's own descriptor grid and the
keypoint cells it selected: every fp16 descriptor row must be bit-identical (same 256-wide tree sum, same rsqrtf,
same rounding) - SURVEY §8 row a9 pinned by the reference itself rather than by a restatement."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, SP_WEIGHTS

pytestmark = pytest.mark.gpu
LIB = os.path.join(ROOT, "oracle", "_ref", "libref_gather.so")


@pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libref_gather.so not built (build() with /root/reference mounted)")
def test_reference_gather_kernel_equals_product_rows():
    import torch

    from oracle import superpoint as osp
    from superslam_b200 import _lib
    from superslam_b200 import frontend as fe

    ref = C.CDLL(LIB)
    ref.ref_launch_gather.restype = C.c_int
    ref.ref_launch_gather.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    imgs = np.load(os.path.join(GOLDEN, "superpoint_ref_small.npz"))["images"]
    b, h, w = imgs.shape
    K = 256
    sp = fe.SuperPoint(SP_WEIGHTS, K)
    feats = sp._extract(list(imgs))
    hc, wc = h // 8, w // 8
    raw = sp.debug_read("scores", (b, hc * 8, wc * 8), np.float32)
    grid = sp.debug_read("grid", (b, hc, wc, 256), np.float16)          # the product stores it cell-major
    checked = 0
    for i, F in enumerate(feats):
        k = osp.nms_select(raw[i], h, w, K, 0.005, 4)
        assert np.array_equal(F.keypoints, k["xy"])                      # same keypoints, same order -> same cells
        n = len(k["xy"])
        if n == 0:
            continue
        chw = torch.from_numpy(np.ascontiguousarray(grid[i].transpose(2, 0, 1))).cuda()   # the reference's layout
        ch = torch.from_numpy(np.ascontiguousarray(k["cell"][:, 0], np.int32)).cuda()
        cw = torch.from_numpy(np.ascontiguousarray(k["cell"][:, 1], np.int32)).cuda()
        out = torch.zeros((n, 256), dtype=torch.float16, device="cuda")
        torch.cuda.synchronize()
        rc = ref.ref_launch_gather(chw.data_ptr(), 256, hc, wc, ch.data_ptr(), cw.data_ptr(), n, out.data_ptr())
        assert rc == 0, f"reference kernel failed: cudaError {rc}"
        theirs = out.cpu().numpy()
        ours = np.zeros((n, 256), np.float32)
        _lib.check(_lib.load().ssb_desc_to_host_f32(0, C.c_void_p(F.descriptors.data), n, 256,
                                                    ours.ctypes.data_as(C.POINTER(C.c_float))))
        assert np.array_equal(ours.astype(np.float16).view(np.uint16), theirs.view(np.uint16))
        norms = np.linalg.norm(theirs.astype(np.float32), axis=1)
        assert np.abs(norms - 1.0).max() < 2e-3
        checked += n
    assert checked > 50


def test_bgr_input_is_converted_like_opencv():
    """IFeatureExtractor::extract on a 3-channel image (src/SuperPoint.cc:387-388: cv::cvtColor BGR2GRAY): the device
    conversion must give exactly the features of the gray image OpenCV itself produces (committed cv2 output)."""
    from superslam_b200 import frontend as fe

    g = np.load(os.path.join(GOLDEN, "imgproc_cv2.npz"))
    sp = fe.SuperPoint(SP_WEIGHTS, 256)
    a = sp.extract(np.ascontiguousarray(g["bgr"]))
    b = sp.extract(np.ascontiguousarray(g["gray"]))
    assert len(b.keypoints) > 10
    assert np.array_equal(a.keypoints, b.keypoints) and np.array_equal(a.responses, b.responses)


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_nethost.so")),
                    reason="oracle/_ref/libref_nethost.so not built (build() with /root/reference mounted)")
@pytest.mark.parametrize("fixture,K,least", [("superpoint_ref_small.npz", 256, 50), ("superpoint_ref_odd.npz", 128, 20)])
def test_reference_select_and_gather_equals_product_features(fixture, K, least):
    """Rows a7-a9 in one piece: the reference's own SuperPoint::select_and_gather (src/SuperPoint.cc:681-750, compiled in
    place, with its own DescriptorPool and gather kernel) fed with the product's heat map (after the graph's NMS) and
    descriptor grid must return the product's keypoints, responses and fp16 descriptor rows bit for bit."""
    import torch

    from oracle import superpoint as osp
    from superslam_b200 import _lib
    from superslam_b200 import frontend as fe
    from test_oracle_ref_superpoint import bind, ref_select

    lib = bind()
    imgs = np.load(os.path.join(GOLDEN, fixture))["images"]
    b, h, w = imgs.shape
    hc, wc = h // 8, w // 8
    sp = fe.SuperPoint(SP_WEIGHTS, K)
    feats = sp._extract(list(imgs))
    raw = sp.debug_read("scores", (b, hc * 8, wc * 8), np.float32)
    grid = sp.debug_read("grid", (b, hc, wc, 256), np.float16)
    total = 0
    for i, F in enumerate(feats):
        chw = torch.from_numpy(np.ascontiguousarray(grid[i].transpose(2, 0, 1))).cuda()   # the reference's CHW binding
        torch.cuda.synchronize()
        got = ref_select(lib, osp.nms(raw[i]), h, w, K, 0.005, 4, grid_dev=C.c_void_p(chw.data_ptr()), want_desc=True)
        assert got["ok"] == 1 and got["n"] == len(F.keypoints)
        assert np.array_equal(got["xy"], F.keypoints) and np.array_equal(got["score"], F.responses)
        ours = np.zeros((got["n"], 256), np.float32)
        if got["n"]:
            _lib.check(_lib.load().ssb_desc_to_host_f32(0, C.c_void_p(F.descriptors.data), got["n"], 256,
                                                        ours.ctypes.data_as(C.POINTER(C.c_float))))
        assert np.array_equal(ours.astype(np.float16).view(np.uint16), got["desc"])
        total += got["n"]
    assert total > least   # sanity only (the 99x131 fixture holds 42 keypoints in all): the equalities above are the test
