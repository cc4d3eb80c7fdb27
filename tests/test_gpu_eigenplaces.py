"""GPU parity of the EigenPlaces path (C-ABI ssb_ep_*) against oracle/eigenplaces.py: preprocess (bit-exact
resize, fp16-rounded normalisation), every ResNet18 stage, the global descriptor, and the device index
against the reference's own retrieval tests."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W_IN, H_IN = 256, 224   # both multiples of 32, not square, small enough for a fast oracle


@pytest.fixture(scope="module")
def setup(tmp_path_factory):
    from oracle import eigenplaces as oep
    from superslam_b200 import frontend as fe
    from superslam_b200.eigenplaces_weights import make_random_weights, save_state_dict

    sd = make_random_weights(11)
    path = str(tmp_path_factory.mktemp("ep") / "eigenplaces.ssbw")
    save_state_dict(sd, path)
    ep = fe.EigenPlaces(path, W_IN, H_IN, max_batch=2, min_score=0.0)
    return ep, oep.load_weights(path), oep


def _nhwc(ep, name, h, w, c, batch=2):
    return ep.debug_read(name, (batch, h, w, c), np.float16).astype(np.float32)


@pytest.mark.parametrize("shape", [(480, 640), (480, 752, 3), (448, 512), (224, 256)])
def test_preprocess_matches_oracle(setup, shape):
    ep, w, oep = setup
    rng = np.random.default_rng(3)
    imgs = [rng.integers(0, 256, size=shape, dtype=np.uint8) for _ in range(2)]
    d = ep.compute_global_descriptors(imgs)
    assert d.shape == (2, 512)
    x0 = _nhwc(ep, "x0", H_IN, W_IN, 4)
    for i, im in enumerate(imgs):
        ref = oep.preprocess(im, W_IN, H_IN).transpose(1, 2, 0)
        # the resize is bit-exact (u8); the only difference is the fp16 rounding of the normalised value
        assert np.array_equal(x0[i, :, :, :3], ref.astype(np.float16).astype(np.float32))
        assert np.all(x0[i, :, :, 3] == 0)


def test_stages_and_descriptor_match_oracle(setup):
    ep, w, oep = setup
    from superslam_b200.synth import synth_pair

    l, r = synth_pair(480, 640, 77)
    d = ep.compute_global_descriptors([l, r])
    taps = {}
    x = np.stack([oep.preprocess(l, W_IN, H_IN), oep.preprocess(r, W_IN, H_IN)])
    ref = oep.forward(w, x, taps)
    geo = {"stem": (H_IN // 2, W_IN // 2, 64), "pool": (H_IN // 4, W_IN // 4, 64), "layer1": (H_IN // 4, W_IN // 4, 64),
           "layer2": (H_IN // 8, W_IN // 8, 128), "layer3": (H_IN // 16, W_IN // 16, 256),
           "layer4": (H_IN // 32, W_IN // 32, 512)}
    for name, (h, wd, c) in geo.items():
        got = _nhwc(ep, name, h, wd, c)
        exp = taps[name].numpy().transpose(0, 2, 3, 1)
        scale = float(np.abs(exp).max())
        err = float(np.abs(got - exp).max())
        assert err <= 6e-3 * scale, f"{name}: max abs err {err:.4g} vs range {scale:.4g}"
    # descriptor: unit rows, within fp16-storage tolerance of the fp32 oracle
    refn = ref / np.linalg.norm(ref, axis=1, keepdims=True)
    assert np.allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-5)
    cos = np.sum(d * refn, axis=1)
    assert np.all(cos > 0.9995), cos
    assert float(np.abs(d - refn).max()) < 3e-3
    # batch of one == row of the batch of two (no cross-image leakage)
    d0 = ep.compute_global_descriptor(l)
    assert np.array_equal(d0[0], d[0])


def test_more_images_than_max_batch(setup):
    ep, w, oep = setup
    rng = np.random.default_rng(9)
    imgs = [rng.integers(0, 256, size=(240, 320), dtype=np.uint8) for _ in range(5)]
    d = ep.compute_global_descriptors(imgs)
    for i in (0, 2, 4):
        assert np.array_equal(d[i], ep.compute_global_descriptor(imgs[i])[0])


def test_invalid_input_returns_empty(setup):
    ep, w, oep = setup
    a = np.zeros((64, 64), np.uint8)
    b = np.zeros((32, 64), np.uint8)
    assert ep.compute_global_descriptors([a, b]).shape == (0, 512)


# ---- the reference's retrieval tests (tests/test_place_recognizer.cc) on the device index ----------
def _desc(dim, seed, jitter=0.0):
    d = np.zeros((1, dim), np.float32)
    d[0, seed % dim] = 1.0
    d[0, (seed + 1) % dim] = 0.5 + jitter
    return d


def _fresh(tmp_path):
    from superslam_b200 import frontend as fe
    from superslam_b200.eigenplaces_weights import make_random_weights, save_state_dict

    path = str(tmp_path / "ep.ssbw")
    if not os.path.exists(path):
        save_state_dict(make_random_weights(11), path)
    return fe.EigenPlaces(path, 64, 64, min_score=0.0)


def test_index_semantics_match_reference_tests(tmp_path):
    from oracle import eigenplaces as oep

    idx = _fresh(tmp_path)
    idx.add(0, _desc(16, 3))
    idx.add(1, _desc(16, 9))
    res = idx.query(_desc(16, 3, 0.01), 0, 5, 0.0)
    assert res and res[0][0] == 0 and res[0][1] > 0.95
    if len(res) > 1:
        assert res[1][1] < res[0][1]

    idx = _fresh(tmp_path)
    for i in range(5):
        idx.add(i, _desc(16, i))
    assert all(k < 3 for k, _ in idx.query(_desc(16, 4), 2, 5, 0.0))

    idx = _fresh(tmp_path)
    ref = oep.CosineDescriptorIndex()
    for i in range(6):
        idx.add(i, _desc(16, i))
        ref.add(i, _desc(16, i))
    assert len(idx.query(_desc(16, 0), 0, 2, -1.0)) <= 2
    assert all(s >= 0.99 for _, s in idx.query(_desc(16, 0), 0, 10, 0.99))
    got, exp = idx.query(_desc(16, 0), 0, 10, -1.0), ref.query(_desc(16, 0), 0, 10, -1.0)
    assert [k for k, _ in got][:2] == [k for k, _ in exp][:2]
    assert np.allclose([s for _, s in got], [s for _, s in exp], atol=1e-6)

    idx = _fresh(tmp_path)
    assert idx.query(_desc(16, 0), 0, 5, 0.0) == []
    idx.add(0, _desc(16, 0))
    assert idx.query(_desc(16, 0), 1, 5, 0.0) == []


def test_index_large_random_matches_oracle(tmp_path):
    from oracle import eigenplaces as oep

    rng = np.random.default_rng(2)
    idx, ref = _fresh(tmp_path), oep.CosineDescriptorIndex()
    rows = rng.standard_normal((700, 512)).astype(np.float32)   # crosses the 256-row growth steps
    for i, r in enumerate(rows):
        idx.add(1000 + i, r)
        ref.add(1000 + i, r)
    assert idx.size() == 700
    q = rows[123] + 0.05 * rng.standard_normal(512).astype(np.float32)
    got, exp = idx.query(q, 30, 10, 0.05), ref.query(q, 30, 10, 0.05)
    assert [k for k, _ in got] == [k for k, _ in exp]
    assert got[0][0] == 1123
    assert np.allclose([s for _, s in got], [s for _, s in exp], atol=1e-5)
