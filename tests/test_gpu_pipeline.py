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
    import parity

    rep = parity.check_matches(pipe.lightglue_debug_read, pipe.kp, lg_weights, out["matches0"][0], out["mscores0"][0],
                               L.keypoints, d0, R.keypoints, d1, w, h, pair=0)
    print(f"K=2048: {rep['differ']} matches0 entries differ (all near-ties), score error {rep['score_err']:.3g}")
    assert out["has_depth"][0].sum() > 50


def test_c5_720p_k4096_dynamic_counts(tmp_path, lg_weights):
    """Config C5: 1280 x 720, K = 4096, pairs with different keypoint counts in one call (dynamic N):
    the batched pipeline must equal the per-pair interface calls exactly, and LightGlue on the sparser
    pair must agree with the oracle."""
    from oracle import lightglue as olg
    from superslam_b200 import frontend as fe
    from superslam_b200.lightglue_weights import save_state_dict
    from superslam_b200.synth import synth_pair

    lgw = str(tmp_path / "lg.ssbw")
    save_state_dict(lg_weights, lgw)
    h, w, K = 720, 1280, 4096
    pairs = [synth_pair(h, w, 1234, 900), synth_pair(h, w, 1235, 3000)]
    pipe = fe.FramePairPipeline(SP_WEIGHTS, lgw, K, w, h, max_pairs=2)
    out = pipe.process([im for p in pairs for im in p])
    cnt = out["count"].tolist()
    assert cnt[2] == K and cnt[3] == K                 # the dense pair saturates K
    assert 256 < cnt[0] < K and 256 < cnt[1] < K       # the sparse one does not: ragged tiles inside one launch
    sp = fe.SuperPoint(SP_WEIGHTS, K)
    lg = fe.LightGlue(lgw, w, h, max_keypoints=K)
    for p, (l, r) in enumerate(pairs):
        L, R = sp.extract_stereo(l, r)
        n0, n1 = cnt[2 * p], cnt[2 * p + 1]
        assert (len(L.keypoints), len(R.keypoints)) == (n0, n1)
        assert np.array_equal(out["xy"][2 * p, :n0], L.keypoints)
        m = lg.match(L.keypoints, L.descriptors, R.keypoints, R.descriptors)
        assert np.array_equal(out["matches0"][p, :n0], m.matches0)
        assert np.array_equal(out["mscores0"][p, :n0], m.mscores0)
        assert np.all(out["matches0"][p, n0:] == -1)
        if p == 0:
            d0, d1 = lg.descriptors_to_host(L.descriptors), lg.descriptors_to_host(R.descriptors)
            import parity

            rep = parity.check_matches(lg.debug_read, lg.kp, lg_weights, m.matches0, m.mscores0, L.keypoints, d0,
                                       R.keypoints, d1, w, h)
            print(f"K=4096, {n0} x {n1}: {rep['differ']} matches0 entries differ (all near-ties), score error {rep['score_err']:.3g}")
    assert out["has_depth"][1].sum() > 100


def test_c4_euroc_size_with_keyframe_descriptor(tmp_path, lg_weights):
    """Config C4: 752 x 480 stereo pairs, K = 1024, plus one EigenPlaces global descriptor per keyframe
    computed from the left image and fed to the loop-closure index."""
    from oracle import eigenplaces as oep
    from superslam_b200 import frontend as fe
    from superslam_b200.eigenplaces_weights import make_random_weights, save_state_dict as save_ep
    from superslam_b200.lightglue_weights import save_state_dict
    from superslam_b200.synth import synth_pair

    lgw, epw = str(tmp_path / "lg.ssbw"), str(tmp_path / "ep.ssbw")
    save_state_dict(lg_weights, lgw)
    save_ep(make_random_weights(11), epw)
    h, w, K = 480, 752, 1024
    pairs = [synth_pair(h, w, 300 + i) for i in range(4)]
    pipe = fe.FramePairPipeline(SP_WEIGHTS, lgw, K, w, h, max_pairs=4)
    out = pipe.process([im for p in pairs for im in p])
    assert out["count"].min() > 500 and out["has_depth"].sum(axis=1).min() > 30
    ep = fe.EigenPlaces(epw, 512, 512, max_batch=4, min_score=0.5)
    lefts = [p[0] for p in pairs]
    d = ep.compute_global_descriptors(lefts)
    ref = oep.compute_global_descriptor(oep.load_weights(epw), lefts[0], 512, 512)
    assert float(np.sum(d[0] * ref[0])) > 0.9995
    for i in range(4):
        ep.add(i, d[i:i + 1])
    # revisiting keyframe 1 (same place, sensor noise): best candidate among all but the last insertion
    noisy = np.clip(lefts[1].astype(np.int16) + np.random.default_rng(0).integers(-3, 4, lefts[1].shape), 0, 255).astype(np.uint8)
    res = ep.query(ep.compute_global_descriptor(noisy), 1, 3)
    assert res and res[0][0] == 1 and res[0][1] > 0.9


def test_streaming_submit_collect_equals_process(tmp_path, lg_weights):
    """ssb_fe_submit / ssb_fe_collect (upload of step i+1 under the kernels of step i, two image buffers,
    graph replay from the third use of a buffer) returns exactly what the synchronous call returns."""
    from superslam_b200 import _lib
    from superslam_b200 import frontend as fe
    from superslam_b200.lightglue_weights import save_state_dict
    from superslam_b200.synth import synth_pair

    lgw = str(tmp_path / "lg.ssbw")
    save_state_dict(lg_weights, lgw)
    h, w, K = 240, 320, 512
    steps = [[im for i in range(2) for im in synth_pair(h, w, 700 + 10 * s + i)] for s in range(7)]
    pipe = fe.FramePairPipeline(SP_WEIGHTS, lgw, K, w, h, max_pairs=2)
    ref = [pipe.process(s) for s in steps]
    got = []
    pipe.submit(steps[0])
    for s in range(1, len(steps)):
        pipe.submit(steps[s])
        got.append(pipe.collect())
    with pytest.raises(Exception):
        pipe.fetch(2)            # a step is still in flight
    got.append(pipe.collect())
    for r, g in zip(ref, got):
        for k in r:
            assert np.array_equal(r[k], g[k], equal_nan=True), k
    # at most two steps in flight
    pipe.submit(steps[0])
    pipe.submit(steps[1])
    with pytest.raises(_lib.SsbError):
        pipe.submit(steps[2])
    pipe.collect()
    pipe.collect()
    out = pipe.process(steps[3])   # the synchronous call works again once everything is collected
    assert np.array_equal(out["matches0"], ref[3]["matches0"])


def test_pipeline_with_a_blank_image_in_the_batch(tmp_path, lg_weights):
    """Device-side counts: one image of a three-pair call has no keypoints at all (attention tiles without keys,
    GEMM tiles without rows, deferred attention epilogues across skipped tiles).  That pair has no matches, the
    other pairs are untouched."""
    from superslam_b200 import frontend as fe
    from superslam_b200.lightglue_weights import save_state_dict
    from superslam_b200.synth import synth_pair

    lgw = str(tmp_path / "lg.ssbw")
    save_state_dict(lg_weights, lgw)
    h, w, K = 240, 320, 512
    pairs = [list(synth_pair(h, w, 50 + i, 100 + 20 * i)) for i in range(3)]
    pipe = fe.FramePairPipeline(SP_WEIGHTS, lgw, K, w, h, max_pairs=3)
    ref = pipe.process([im for p in pairs for im in p])
    ref = {k: v.copy() for k, v in ref.items()}
    blank = np.full((h, w), 128, np.uint8)
    for which in (0, 1):                       # blank left image, then blank right image, of the middle pair
        mod = [list(p) for p in pairs]
        mod[1][which] = blank
        for _ in range(3):                     # eager, capture, replay
            out = pipe.process([im for p in mod for im in p])
        assert out["count"][2 + which] == 0 and out["count"][2 + (1 - which)] == ref["count"][2 + (1 - which)]
        assert (out["matches0"][1] == -1).all() and not out["has_depth"][1].any()
        for p in (0, 2):
            n0 = ref["count"][2 * p]
            assert np.array_equal(out["matches0"][p, :n0], ref["matches0"][p, :n0])
            assert np.array_equal(out["mscores0"][p, :n0], ref["mscores0"][p, :n0])
            assert np.array_equal(out["has_depth"][p], ref["has_depth"][p])


def test_cached_graphs_survive_workspace_reallocation(tmp_path, lg_weights):
    """One pipeline, sizes and pair counts revisited after a change.  Captured CUDA graphs hold SuperPoint's activation
    pointers and tensor maps; a larger batch or another image size reallocates them (SuperPoint::ensure_shape), so every
    cached graph must be dropped then - a stale replay would run on freed memory.  Each visit must reproduce the first
    result for that input exactly (calls 2 and 3 of a key are the capture and the first replay)."""
    from superslam_b200 import frontend as fe
    from superslam_b200.lightglue_weights import save_state_dict
    from superslam_b200.synth import synth_pair

    lgw = str(tmp_path / "lg.ssbw")
    save_state_dict(lg_weights, lgw)
    K = 256
    big = [im for i in range(4) for im in synth_pair(240, 320, 7 + i, 60)]
    small = [im for i in range(2) for im in synth_pair(120, 160, 70 + i, 40)]
    pipe = fe.FramePairPipeline(SP_WEIGHTS, lgw, K, 320, 240, max_pairs=4)
    keys = ("count", "xy", "score", "matches0", "mscores0", "has_depth")

    def snap(o):
        return {k: o[k].copy() for k in keys}

    def same(a, b):
        return all(np.array_equal(a[k], b[k]) for k in keys)

    first = {}
    # host path (shared upload buffer): 1 pair x3 (eager, capture, replay) -> other size -> more pairs -> revisit
    plan = [("b1", big[:2])] * 3 + [("s1", small[:2])] * 3 + [("b1", big[:2])] * 3 + [("b4", big)] * 3 + \
           [("b1", big[:2])] * 3 + [("s2", small)] * 3 + [("b4", big)] * 2 + [("s1", small[:2])] * 2
    for name, imgs in plan:
        got = snap(pipe.process(imgs))
        if name in first:
            assert same(got, first[name]), f"result for {name} changed after a workspace reallocation"
        else:
            first[name] = got
    # caller-owned device buffers: pairs = 1 twice, pairs = 4 once, pairs = 1 again (the advisor's sequence)
    d1 = pipe.upload(big[:2])
    d4 = pipe.upload(big)
    for dev, pairs, name in [(d1, 1, "b1"), (d1, 1, "b1"), (d4, 4, "b4"), (d1, 1, "b1"), (d1, 1, "b1"), (d4, 4, "b4")]:
        pipe.enqueue_device(dev, pairs, 240, 320)
        assert same(snap(pipe.fetch(pairs)), first[name])
    assert int(first["b4"]["count"].min()) > 20


def test_tracking_chain_equals_two_interface_calls(tmp_path, lg_weights):
    """the live pipeline makes TWO LightGlue calls per frame - the stereo match and the tracking match of
    the last keyframe's left features against the current left features (src/VoEstimator.cc:240-246) - and promotes the
    frame to keyframe on the host's decision (:327).  With tracking enabled the pair pipeline keeps each stream's
    keyframe on the device and chains both matches in one graph; it must give exactly what the interface calls give:
    matcher.match(kf.keypoints, kf.descriptors, L.keypoints, L.descriptors), plus the depth test of :253-256."""
    from superslam_b200 import frontend as fe
    from superslam_b200.lightglue_weights import save_state_dict
    from superslam_b200.synth import synth_pair

    lgw = str(tmp_path / "lg.ssbw")
    save_state_dict(lg_weights, lgw)
    h, w, K, S = 240, 320, 512, 2
    sp = fe.SuperPoint(SP_WEIGHTS, K, num_slots=16)
    lg = fe.LightGlue(lgw, w, h, max_keypoints=K)
    front = fe.StereoFrontEnd(sp, lg)
    pipe = fe.FramePairPipeline(SP_WEIGHTS, lgw, K, w, h, max_pairs=S)
    pipe.enable_tracking(True)
    # two streams; frame t of stream s = the stream's scene shifted by 3 t pixels (camera motion)
    base = [synth_pair(h, w, 300 + s, 120) for s in range(S)]
    frames = [[tuple(np.roll(im, -3 * t, axis=1) for im in base[s]) for s in range(S)] for t in range(4)]
    kf = [None] * S        # interface-call side: (Features of the keyframe's left image, has_depth)
    promote_plan = [np.array([1, 1], np.uint8), np.array([1, 0], np.uint8), np.array([0, 0], np.uint8), None]
    for t in range(4):
        flat = [im for s in range(S) for im in frames[t][s]]
        out = pipe.process(flat)
        trk = pipe.tracking_results(S)
        cur = []
        for s in range(S):
            L, R = sp.extract_stereo(*frames[t][s])
            m = lg.match(L.keypoints, L.descriptors, R.keypoints, R.descriptors)
            frame = front.process(*frames[t][s])
            n0 = len(L.keypoints)
            assert np.array_equal(out["matches0"][s, :n0], m.matches0)
            cur.append((L, frame.has_depth.astype(np.uint8)))
            if kf[s] is None:
                assert trk["keyframe_count"][s] == 0 and np.all(trk["track_matches0"][s] == -1)
                assert not trk["track_usable"][s].any()
                continue
            KF, kf_hd = kf[s]
            nk = len(KF.keypoints)
            assert trk["keyframe_count"][s] == nk
            tm = lg.match(KF.keypoints, KF.descriptors, L.keypoints, L.descriptors)
            assert np.array_equal(trk["track_matches0"][s, :nk], tm.matches0)
            assert np.array_equal(trk["track_mscores0"][s, :nk], tm.mscores0)
            usable = np.zeros(K, np.uint8)
            ok = tm.matches0 >= 0
            usable[:nk][ok] = kf_hd[:nk][ok] & cur[s][1][tm.matches0[ok]]
            assert np.array_equal(trk["track_usable"][s], usable)
            if t == 1:
                assert ok.sum() > 20   # the scene moved by 3 px: most keyframe features are found again
        mask = promote_plan[t]
        if mask is not None:
            pipe.promote_keyframes(S, mask)
            for s in range(S):
                if mask[s]:
                    kf[s] = cur[s]
    # tracking off again: the stereo results do not change, tracking results are refused
    pipe.enable_tracking(False)
    again = pipe.process([im for s in range(S) for im in frames[3][s]])
    assert np.array_equal(again["matches0"], out["matches0"])
    with pytest.raises(Exception):
        pipe.tracking_results(S)


def test_c2_end_to_end_against_the_oracle(tmp_path, lg_weights, sp_weights):
    """ONE 640 x 480 pair at K = 1024 through FramePairPipeline against the oracle's composition
    extract -> match -> dmatches -> stereo_postfilter on the same images (the composition tests/test_oracle_ref_e2e.py
    pins bit for bit to the reference's own StereoFrontEnd::process).  Keypoints: identical up to near-ties of the
    reference heat map; descriptors within 1e-3; matches0 on the pipeline's own features: identical up to near-ties of
    the oracle's assignment scores; the post-filter exact on those matches; and the frame's depth flags equal to the
    oracle's wherever the two sides agree on the keypoints and matches involved."""
    import parity
    from oracle import frontend as ofe
    from oracle import superpoint as osp
    from superslam_b200 import frontend as fe
    from superslam_b200.lightglue_weights import save_state_dict
    from superslam_b200.synth import synth_pair

    lgw = str(tmp_path / "lg.ssbw")
    save_state_dict(lg_weights, lgw)
    h, w, K = 480, 640, 1024
    l, r = synth_pair(h, w, 4321)
    pipe = fe.FramePairPipeline(SP_WEIGHTS, lgw, K, w, h, max_pairs=1)
    out = pipe.process([l, r])
    n0, n1 = int(out["count"][0]), int(out["count"][1])
    # ---- oracle, end to end
    ref = osp.extract(np.stack([l, r]), sp_weights, K)
    raw_gpu = pipe.superpoint_debug_read("scores", (2, h, w), np.float32)
    for i, n in enumerate((n0, n1)):
        rep = parity.check_keypoints(raw_gpu[i], ref[i]["raw_map"], out["xy"][i, :n], ref[i]["xy"], h, w, K)
        print(f"image {i}: {rep['differ']} keypoints differ of {rep['n_ref']} (all near-ties)")
    # descriptors of the common keypoints (the pipeline's rows through the interface's descriptors_to_host path)
    sp, lg = fe.SuperPoint(SP_WEIGHTS, K), fe.LightGlue(lgw, w, h, max_keypoints=K)
    L, R = sp.extract_stereo(l, r)
    assert np.array_equal(L.keypoints, out["xy"][0, :n0]) and np.array_equal(R.keypoints, out["xy"][1, :n1])
    d = [lg.descriptors_to_host(L.descriptors), lg.descriptors_to_host(R.descriptors)]
    for i, F in enumerate((L, R)):
        ia = {tuple(x): j for j, x in enumerate(ref[i]["xy"].tolist())}
        common = [(j, ia[tuple(x)]) for j, x in enumerate(F.keypoints.tolist()) if tuple(x) in ia]
        gj, rj = map(list, zip(*common))
        assert np.abs(d[i][gj] - ref[i]["desc"][rj].astype(np.float32)).max() < parity.DESC_TOL
    # ---- LightGlue on the pipeline's own features against the oracle on the same features
    rep = parity.check_matches(pipe.lightglue_debug_read, pipe.kp, lg_weights, out["matches0"][0, :n0], out["mscores0"][0, :n0],
                               L.keypoints, d[0], R.keypoints, d[1], w, h)
    print(f"matches0: {rep['differ']} of {n0} differ (all near-ties), score error {rep['score_err']:.3g}, {rep['oracle_matches']} matches")
    # ---- post-filter: exact on the pipeline's matches (src/StereoFrontEnd.cc:35-47)
    q, t, _ = ofe.dmatches(out["matches0"][0, :n0], out["mscores0"][0, :n0])
    st, hd = ofe.stereo_postfilter(L.keypoints, R.keypoints, q, t)
    assert np.array_equal(hd.astype(np.uint8), out["has_depth"][0, :n0])
    ur = out["stereo_ur"][0, :n0]
    assert np.array_equal(np.isnan(ur), np.isnan(st[:, 1])) and np.array_equal(ur[hd == 1].astype(np.float64), st[hd == 1, 1])
    assert hd.sum() > 100
    # ---- and against the oracle's own frame wherever both sides hold the same left keypoint matched to the same right one
    q2, t2, _ = ofe.dmatches(rep["om0"], rep["oms0"])
    _, hd_or = ofe.stereo_postfilter(L.keypoints, R.keypoints, q2, t2)
    same = out["matches0"][0, :n0] == rep["om0"]
    assert np.array_equal(hd_or[same].astype(np.uint8), out["has_depth"][0, :n0][same])
