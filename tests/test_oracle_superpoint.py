"""Pin the SuperPoint oracle against fixtures produced by the reference's own torch module
(tests/golden/make_golden.py) and check the restated host logic on hand-made cases."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import superpoint as osp


def _g(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.mark.parametrize("name", ["superpoint_ref_small.npz", "superpoint_ref_odd.npz"])
def test_dense_forward_matches_reference_bit_exact(sp_weights, name):
    g = _g(name)
    s, d, _ = osp.dense_forward(g["images"], sp_weights)
    assert s.shape == g["scores"].shape and d.shape == g["grid"].shape
    # same torch ops in the same order on the same weights: identical bits
    assert np.array_equal(s, g["scores"])
    assert np.array_equal(d, g["grid"])


def test_odd_size_score_map_is_narrower_than_image(sp_weights):
    g = _g("superpoint_ref_odd.npz")
    assert g["images"].shape[1:] == (99, 131)
    assert g["scores"].shape[1:] == (96, 128)  # three floor-halvings, x8
    k = osp.extract(g["images"], sp_weights, 64)[0]
    # x = w * (131/128): the reference's scale is input/score, not 1 (SuperPoint.cc:708-709)
    sx = np.float32(131) / np.float32(128)
    assert np.array_equal(k["xy"][:, 0], k["hw"][:, 1].astype(np.float32) * sx)


def test_c2_candidates_and_gather_match_reference(sp_weights):
    from superslam_b200.synth import synth_pair

    g = _g("superpoint_ref_c2.npz")
    l, r = synth_pair(480, 640, 1234)
    res = osp.extract(np.stack([l, r]), sp_weights, 1024)
    for i in range(2):
        hw, sc, rows = g[f"hw{i}"], g[f"score{i}"], g[f"rows{i}"]
        assert len(sc) > 1024  # workload saturates max_keypoints
        # expected: candidates sorted by (score desc, h desc, w desc), first 1024
        order = sorted(range(len(sc)), key=lambda j: (-float(sc[j]), -int(hw[j, 0]), -int(hw[j, 1])))[:1024]
        assert np.array_equal(res[i]["hw"], hw[order])
        assert np.array_equal(res[i]["score"], sc[order])
        assert np.array_equal(res[i]["xy"], hw[order][:, ::-1].astype(np.float32))
        # gather: recompute from the reference grid rows with an independent (float64) norm
        v = rows[order].astype(np.float64)
        expect = v / np.sqrt((v * v).sum(1, keepdims=True) + 1e-12)
        got = res[i]["desc"].astype(np.float64)
        assert np.abs(got - expect).max() <= 2.0 ** -11  # one fp16 rounding of values < 1
        assert np.abs(np.linalg.norm(got, axis=1) - 1).max() < 2e-3


def test_kitti_shape_and_candidates(sp_weights):
    from superslam_b200.synth import synth_image

    g = _g("superpoint_ref_kitti.npz")
    assert tuple(g["score_shape"]) == (376, 1240) and tuple(g["grid_shape"]) == (256, 47, 155)
    img = synth_image(376, 1241, 1234)[None]
    k = osp.extract(img, sp_weights, 2048)[0]
    hw, sc = g["hw"], g["score"]
    order = sorted(range(len(sc)), key=lambda j: (-float(sc[j]), -int(hw[j, 0]), -int(hw[j, 1])))[:2048]
    assert np.array_equal(k["hw"], hw[order]) and np.array_equal(k["score"], sc[order])
    assert np.array_equal(k["xy"][:, 0], hw[order][:, 1].astype(np.float32) * (np.float32(1241) / np.float32(1240)))


def test_select_tie_break_threshold_and_borders():
    s = np.zeros((32, 40), np.float32)
    s[10, 10] = 0.5
    s[20, 30] = 0.5   # tie: larger row first
    s[20, 12] = 0.5   # tie on row: larger col first
    s[3, 20] = 0.9    # inside 4-px border -> dropped
    s[15, 36] = 0.9   # w >= W-4 -> dropped
    s[15, 35] = 0.25
    s[16, 16] = 0.005  # not strictly greater than the threshold (0.005f < 0.005 as double? check below)
    k = osp.select_keypoints(s, 32, 40, 10, 0.005, 4, 4, 5)
    assert k["hw"].tolist()[:4] == [[20, 30], [20, 12], [10, 10], [15, 35]]
    # float(0.005) = 0.004999999888... < 0.005 (double): rejected by the promoted compare
    assert [16, 16] not in k["hw"].tolist()
    assert k["cell"].tolist()[0] == [2, 3] and k["cell"].tolist()[3] == [1, 4]
    k2 = osp.select_keypoints(s, 32, 40, 2, 0.005, 4, 4, 5)
    assert len(k2["score"]) == 2
    empty = osp.select_keypoints(np.zeros((32, 40), np.float32), 32, 40, 10, 0.005, 4, 4, 5)
    assert empty["xy"].shape == (0, 2) and empty["cell"].shape == (0, 2)


def test_gather_tree_sum_and_fp16_rounding():
    rng = np.random.default_rng(0)
    grid = rng.normal(size=(256, 3, 4)).astype(np.float16)
    cell = np.array([[0, 0], [2, 3], [1, 2]], np.int32)
    out = osp.gather_normalize(grid, cell)
    assert out.dtype == np.float16 and out.shape == (3, 256)
    v = grid[:, 2, 3].astype(np.float64)
    assert np.abs(out[1].astype(np.float64) - v / np.linalg.norm(v)).max() < 1e-3
    # zero vector: rsqrt(0 + 1e-12) * 0 = 0, no NaN
    z = osp.gather_normalize(np.zeros((256, 1, 1), np.float16), np.zeros((1, 2), np.int32))
    assert np.array_equal(z, np.zeros((1, 256), np.float16))
    assert osp.gather_normalize(grid, np.zeros((0, 2), np.int32)).shape == (0, 256)


def test_fp16_storage_model_stays_close_to_fp32(sp_weights):
    g = _g("superpoint_ref_small.npz")
    a = osp.extract(g["images"], sp_weights, 256)
    b = osp.extract(g["images"], sp_weights, 256, fp16_storage=True)
    for x, y in zip(a, b):
        # the fp16-storage restatement selects the same keypoints up to near-ties (the rule of tests/parity.py)
        rep = osp.keypoint_disagreements(x["raw_map"], y["raw_map"], x["hw"], y["hw"], 256)
        assert not rep["unexplained"], rep["unexplained"][:3]
        assert np.abs(x["grid_f16"].astype(np.float32) - y["grid_f16"].astype(np.float32)).max() < 2e-3


def test_keypoint_disagreements_separate_near_tie_flips_from_defects():
    """The index-parity rule the GPU tests assert (oracle/superpoint.py::keypoint_disagreements): keypoint sets selected from
    two heat maps that differ by a small error may only differ at decisions whose margin in the oracle's map is below twice
    the locally measured error; a planted defect (a far-from-tie keypoint dropped, a non-maximum added) is reported."""
    from oracle import superpoint as osp

    rng = np.random.default_rng(0)
    H, W, K = 96, 128, 100
    raw = (rng.random((H, W)) ** 6 * 0.5).astype(np.float32)            # a few hundred local maxima above 0.005
    noise = (rng.normal(0, 2e-3, raw.shape)).astype(np.float32)
    other = np.clip(raw + noise, 0, None).astype(np.float32)
    a = osp.nms_select(raw, H, W, K, 0.005, 4)
    b = osp.nms_select(other, H, W, K, 0.005, 4)
    assert len(a["hw"]) == K
    rep = osp.keypoint_disagreements(raw, other, a["hw"], b["hw"], K)
    assert rep["differ"] > 0 and not rep["unexplained"], rep["unexplained"][:3]     # only near-tie flips
    assert all(r["bound"] <= 2 * np.abs(noise).max() for r in rep["rows"])
    # defect 1: the strongest keypoint is missing from the other side's set
    without = b["hw"][~np.all(b["hw"] == a["hw"][0], axis=1)]
    assert len(without) == len(b["hw"]) - 1
    rep = osp.keypoint_disagreements(raw, other, a["hw"], without, K)
    bad = [r for r in rep["unexplained"] if r["hw"] == tuple(a["hw"][0])]
    assert len(bad) == 1 and bad[0]["oracle_selected"] and bad[0]["margin"] > 3 * bad[0]["bound"]
    # defect 2: a pixel right next to a strong maximum (not a local maximum itself) is reported as a keypoint
    h0, w0 = a["hw"][0]
    fake = np.concatenate([b["hw"], [[h0, w0 + 1]]])
    rep = osp.keypoint_disagreements(raw, other, a["hw"], fake, K)
    assert any(r["hw"] == (h0, w0 + 1) and r["why"] == "nms" for r in rep["unexplained"])
    # defect 3: a keypoint inside the border band
    rep = osp.keypoint_disagreements(raw, other, a["hw"], np.concatenate([b["hw"], [[1, 50]]]), K)
    assert any(r["hw"] == (1, 50) and r["margin"] == np.inf for r in rep["unexplained"])
    # identical sets: nothing to explain
    rep = osp.keypoint_disagreements(raw, raw, a["hw"], a["hw"], K)
    assert rep["differ"] == 0 and not rep["unexplained"]
