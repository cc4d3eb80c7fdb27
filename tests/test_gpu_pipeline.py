"""The chained frame-pair pipeline (ssb_fe_*) must give exactly what the reference-shaped calls give
(extract_stereo + match + StereoFrontEnd filter), for one and for several pairs per call."""
import numpy as np
import pytest

from conftest import SP_WEIGHTS

pytestmark = pytest.mark.gpu


def test_pipeline_equals_interface_calls(tmp_path, lg_weights):
    from superslam_b200 import frontend as fe
    from superslam_b200.lightglue_weights import save_state_dict
    from superslam_b200.synth import synth_pair

    lgw = str(tmp_path / "lg.ssbw")
    save_state_dict(lg_weights, lgw)
    h, w, K = 240, 320, 512
    pairs = [synth_pair(h, w, 50 + i, 100 + 20 * i) for i in range(3)]
    sp = fe.SuperPoint(SP_WEIGHTS, K)
    lg = fe.LightGlue(lgw, w, h, max_keypoints=K)
    front = fe.StereoFrontEnd(sp, lg)
    pipe = fe.FramePairPipeline(SP_WEIGHTS, lgw, K, w, h, max_pairs=3)
    flat = [im for p in pairs for im in p]
    out = pipe.process(flat)
    one = pipe.process(flat[:2])
    for p, (l, r) in enumerate(pairs):
        L, R = sp.extract_stereo(l, r)
        m = lg.match(L.keypoints, L.descriptors, R.keypoints, R.descriptors)
        frame = front.process(l, r)
        n0, n1 = out["count"][2 * p], out["count"][2 * p + 1]
        assert (n0, n1) == (len(L.keypoints), len(R.keypoints))
        assert np.array_equal(out["xy"][2 * p, :n0], L.keypoints) and np.array_equal(out["score"][2 * p + 1, :n1], R.responses)
        assert np.array_equal(out["matches0"][p, :n0], m.matches0)
        assert np.array_equal(out["mscores0"][p, :n0], m.mscores0)
        assert np.array_equal(out["has_depth"][p, :n0].astype(np.int8), frame.has_depth)
        ur = out["stereo_ur"][p, :n0]
        assert np.array_equal(np.isnan(ur), np.isnan(frame.stereo[:, 1]))
        assert np.array_equal(ur[~np.isnan(ur)].astype(np.float64), frame.stereo[~np.isnan(ur), 1])
    assert np.array_equal(one["matches0"][0], out["matches0"][0]) and np.array_equal(one["xy"][:2], out["xy"][:2])
    # true matches of the synthetic pair have disparity 12 and pass the filter
    assert out["has_depth"].sum() > 30


def test_kitti_size_pipeline_k2048(tmp_path, lg_weights):
    """1241 x 376 (odd width: score map 1240 wide, x scale 1241/1240), K = 2048, two pairs per call."""
    from oracle import lightglue as olg
    from superslam_b200 import frontend as fe
    from superslam_b200.lightglue_weights import save_state_dict
    from superslam_b200.synth import synth_pair

    lgw = str(tmp_path / "lg.ssbw")
    save_state_dict(lg_weights, lgw)
    h, w, K = 376, 1241, 2048
    pairs = [synth_pair(h, w, 1234 + i) for i in range(2)]
    pipe = fe.FramePairPipeline(SP_WEIGHTS, lgw, K, w, h, max_pairs=2)
    out = pipe.process([im for p in pairs for im in p])
    assert out["count"].tolist() == [K] * 4
    sx = np.float32(w) / np.float32(1240)
    xs = out["xy"][0, :, 0]
    assert np.array_equal(xs, np.round(xs / sx).astype(np.float32) * sx)      # x = w * (1241/1240)
    assert np.all(np.diff(out["score"][0]) <= 0)
    # LightGlue on the pipeline's own keypoints / descriptors == oracle on the same inputs
    sp = fe.SuperPoint(SP_WEIGHTS, K)
    lg = fe.LightGlue(lgw, w, h, max_keypoints=K)
    L, R = sp.extract_stereo(*pairs[0])
    assert np.array_equal(L.keypoints, out["xy"][0]) and np.array_equal(R.keypoints, out["xy"][1])
    d0, d1 = lg.descriptors_to_host(L.descriptors), lg.descriptors_to_host(R.descriptors)
    om0, _ = olg.match(lg_weights, olg.normalize_keypoints(L.keypoints, w, h), d0,
                       olg.normalize_keypoints(R.keypoints, w, h), d1)
    assert (out["matches0"][0] == om0).mean() >= 0.98
    assert out["has_depth"][0].sum() > 50
