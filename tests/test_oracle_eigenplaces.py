"""Pin the EigenPlaces oracle: the resize port against cv2.resize (the routine the reference calls,
src/EigenPlaces.cc:129), the ResNet18 trunk against torchvision with shared weights, and the retrieval
classes against the reference's own unit tests (tests/test_place_recognizer.cc, re-expressed)."""
import numpy as np
import pytest
import torch

from oracle import eigenplaces as oep
from superslam_b200.eigenplaces_weights import make_random_weights


@pytest.mark.parametrize("shape,dst", [((480, 640), (512, 512)), ((480, 752), (512, 512)), ((376, 1241), (512, 512)),
                                       ((100, 60), (37, 91)), ((64, 64), (32, 32)), ((33, 47), (99, 141))])
@pytest.mark.parametrize("cn", [1, 3])
def test_resize_port_matches_cv2_bit_exact(shape, dst, cn):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, size=shape + ((cn,) if cn == 3 else ()), dtype=np.uint8)
    ref = cv2.resize(img, dst)   # dsize = (width, height), default INTER_LINEAR
    got = oep.resize_linear_u8(img, dst[0], dst[1])
    assert got.shape == ref.shape
    assert np.array_equal(got, ref)


def test_preprocess_layout_and_normalisation():
    img = np.full((48, 64), 128, np.uint8)
    x = oep.preprocess(img, 32, 32)
    assert x.shape == (3, 32, 32) and x.dtype == np.float32
    v = np.float32(128) * np.float32(1.0 / 255.0)
    for c in range(3):
        assert np.allclose(x[c], (v - oep.MEAN[c]) / oep.STD[c], atol=1e-7)
    # BGR input: channel order reversed to RGB (EigenPlaces.cc:127)
    bgr = np.zeros((48, 64, 3), np.uint8)
    bgr[..., 0] = 255   # blue
    x = oep.preprocess(bgr, 32, 32)
    assert x[2].mean() > 2.0 and x[0].mean() < -2.0


def test_backbone_matches_torchvision_resnet18():
    tv = pytest.importorskip("torchvision")
    w = make_random_weights(3)
    m = tv.models.resnet18(weights=None).eval()
    bb = torch.nn.Sequential(*list(m.children())[:-2])   # the model's own construction (eigenplaces hub code)
    sd = {k[len("backbone."):]: v for k, v in w.items() if k.startswith("backbone.")}
    missing = bb.load_state_dict(sd, strict=False)
    assert all(k.endswith("num_batches_tracked") for k in missing.missing_keys) and not missing.unexpected_keys
    x = torch.randn(2, 3, 96, 128, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        ref = bb(x)
        got = oep.backbone(x, w)
    assert got.shape == ref.shape == (2, 512, 3, 4)
    assert torch.allclose(got, ref, rtol=1e-5, atol=1e-5)


def test_descriptor_is_unit_norm_and_deterministic():
    w = make_random_weights(3)
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, size=(120, 160), dtype=np.uint8)
    d0 = oep.compute_global_descriptor(w, img, 128, 128)
    d1 = oep.compute_global_descriptor(w, img, 128, 128)
    assert d0.shape == (1, 512) and np.array_equal(d0, d1)
    assert abs(float(np.linalg.norm(d0)) - 1.0) < 1e-5


# ---- tests/test_place_recognizer.cc re-expressed -------------------------------------------------
def _desc(dim, seed, jitter=0.0):
    d = np.zeros((1, dim), np.float32)
    d[0, seed % dim] = 1.0
    d[0, (seed + 1) % dim] = 0.5 + jitter
    return d


def test_index_ranks_near_duplicate_above_distinct():
    idx = oep.CosineDescriptorIndex()
    idx.add(0, _desc(16, 3))
    idx.add(1, _desc(16, 9))
    res = idx.query(_desc(16, 3, 0.01), 0, 5, 0.0)
    assert res and res[0][0] == 0 and res[0][1] > 0.95
    if len(res) > 1:
        assert res[1][1] < res[0][1]


def test_index_exclude_recent_skips_temporal_neighbours():
    idx = oep.CosineDescriptorIndex()
    for i in range(5):
        idx.add(i, _desc(16, i))
    res = idx.query(_desc(16, 4), 2, 5, 0.0)
    assert all(kid < 3 for kid, _ in res)


def test_index_topk_and_min_score_gate():
    idx = oep.CosineDescriptorIndex()
    for i in range(6):
        idx.add(i, _desc(16, i))
    assert len(idx.query(_desc(16, 0), 0, 2, -1.0)) <= 2
    assert all(s >= 0.99 for _, s in idx.query(_desc(16, 0), 0, 10, 0.99))


def test_index_empty_or_all_excluded_returns_nothing():
    idx = oep.CosineDescriptorIndex()
    assert idx.query(_desc(16, 0), 0, 5, 0.0) == []
    idx.add(0, _desc(16, 0))
    assert idx.query(_desc(16, 0), 1, 5, 0.0) == []


def test_voter_requires_consecutive_consistent_votes():
    v = oep.TemporalConsistencyVoter(3, 2)
    a, b = (10, 0.9), (11, 0.9)
    assert not v.vote(a)
    assert not v.vote(b)
    assert v.vote(a)


def test_voter_resets_on_gap_or_inconsistency():
    v = oep.TemporalConsistencyVoter(2, 1)
    a, far = (10, 0.9), (99, 0.9)
    assert not v.vote(a)
    assert not v.vote(None)
    assert not v.vote(a)
    assert not v.vote(far)
    assert v.vote(far)
