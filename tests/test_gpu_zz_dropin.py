"""Drop-in check on the GPU: the REFERENCE's own StereoFrontEnd::process (src/StereoFrontEnd.cc, compiled in place into
oracle/_ref/libdropin.so) on top of the C++ adapter classes and the real library must give exactly what the Python
mirror of the same interfaces gives (tests/test_gpu_lightglue.py pins that one against the oracle): keypoints, stereo
points, depth flags; then the reference's other calls through IFeatureMatcher - keyframe descriptors_to_host, the
tracking match on device descriptors, loop verification with host descriptors on a cloned context.
(The adapter's marshalling itself is checked on the CPU against a C-ABI test double, tests/test_dropin_adapter.py.)"""
import os

import numpy as np
import pytest

from conftest import SP_WEIGHTS
from test_dropin_adapter import REAL, Harness, bind

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not os.path.exists(REAL), reason="oracle/_ref/libdropin.so not built (build() with /root/reference mounted)")
def test_reference_stereo_frontend_over_the_adapter_on_the_gpu(tmp_path, lg_weights):
    from superslam_b200 import frontend as fe
    from superslam_b200.lightglue_weights import save_state_dict
    from superslam_b200.synth import synth_pair

    lgw = str(tmp_path / "lg.ssbw")
    save_state_dict(lg_weights, lgw)
    h, w, K = 240, 320, 512
    hs = Harness(bind(REAL), SP_WEIGHTS, lgw, w, h, max_kp=K)
    assert hs.status == 7, "SuperPointB200 / LightGlueB200 / cloned context failed to initialize"
    sp = fe.SuperPoint(SP_WEIGHTS, K)
    lg = fe.LightGlue(lgw, w, h, max_keypoints=K)
    mirror = fe.StereoFrontEnd(sp, lg)

    def padded(img):   # cv::Mat with step > cols: the adapter passes step[0] through
        buf = np.zeros((img.shape[0], img.shape[1] + 32), np.uint8)
        buf[:, :img.shape[1]] = img
        return buf[:, :img.shape[1]]

    l, r = synth_pair(h, w, 77, 120)
    out = hs.process(padded(l), padded(r), 1.0)
    frame = mirror.process(l, r, 1.0)
    assert out["n"] == len(frame.keypoints_left) > 100
    assert np.array_equal(out["xy"], frame.keypoints_left)
    assert np.all(out["size_angle"] == np.float32([1.0, -1.0])) and np.all(np.diff(out["response"]) <= 0)
    assert np.array_equal(out["has_depth"], frame.has_depth) and out["has_depth"].sum() > 20
    assert np.array_equal(out["stereo"], frame.stereo, equal_nan=True)
    assert out["desc"]["count"] == out["n"] and out["desc"]["dim"] == 256 and out["desc"]["resident"]

    n, ixy, iresp, idesc = hs.infer(padded(l))             # SuperPoint::infer, the demo programs' host path
    assert n == out["n"] and np.array_equal(ixy, out["xy"]) and np.array_equal(iresp, out["response"])
    assert np.array_equal(idesc, lg.descriptors_to_host(frame.descriptors_left))
    assert np.abs(np.linalg.norm(idesc, axis=1) - 1).max() < 2e-3
    ok, pxy, presp, pdesc = sp.infer(l)                    # the Python mirror of the same call
    assert ok and np.array_equal(pxy, ixy) and np.array_equal(presp, iresp) and np.array_equal(pdesc, idesc)

    rows, kf = hs.promote_keyframe()                       # last_keyframe_ = frame; descriptors_to_host
    host0 = lg.descriptors_to_host(frame.descriptors_left)
    assert rows == out["n"] and np.array_equal(kf, host0)

    l2, r2 = np.roll(l, -3, axis=1), np.roll(r, -3, axis=1)   # the next frame of the stream
    out2 = hs.process(padded(l2), padded(r2), 2.0)
    frame2 = mirror.process(l2, r2, 2.0)
    assert np.array_equal(out2["xy"], frame2.keypoints_left) and np.array_equal(out2["has_depth"], frame2.has_depth)
    q, t, d = hs.track()                                   # keyframe <-> frame on device descriptors
    m = lg.match(frame.keypoints_left, frame.descriptors_left, frame2.keypoints_left, frame2.descriptors_left)
    assert len(q) > 0 and np.array_equal(q, m.query) and np.array_equal(t, m.train)
    assert np.allclose(d, m.distance, atol=1e-6)
    q2, t2, d2 = hs.verify()                               # host descriptors, cloned context (loop matcher)
    m2 = lg.shared_context().match(frame.keypoints_left, host0, frame2.keypoints_left,
                                   lg.descriptors_to_host(frame2.descriptors_left))
    assert np.array_equal(q2, m2.query) and np.array_equal(t2, m2.train) and np.allclose(d2, m2.distance, atol=1e-6)
    hs.release_frames()
    hs.close()
