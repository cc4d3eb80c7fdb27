"""The parity rules the GPU tests assert against the fp32 oracle (north_star: "exactly on keypoint / match indices").

The CUDA path stores activations in fp16 (the reference's own `--fp16` engine class), so a decision that hangs on a
near-tie in the fp32 oracle can legitimately fall the other way.  What the tests demand instead of a percentage:

  1. the SCORES the decisions are taken on agree with the oracle's within a stated tolerance, and
  2. EVERY index that differs is a decision whose margin in the oracle's own scores is below twice the error actually
     measured in this run on the scores involved (each of the two competitors can move by that error) -
     oracle/lightglue.py::disagreement_report / explained_by_score_error for matches0,
     oracle/superpoint.py::keypoint_disagreements for the keypoint set.

A wrong index with a clear margin fails whatever the overall agreement is; a flip rate is only reported.
"""
import numpy as np

# log-assignment score error allowed on the entries that take part in a decision, relative to the magnitude of the
# similarity logits the scores are made of (fp16 operands, fp32 accumulation, 18 blocks): measured 3e-4 (B200, C2)
SCORE_TOL_REL = 1e-3
# heat-map error allowed anywhere on the map (softmax outputs in [0, 1], fp16 storage through ten layers); measured 1.6e-3 - 4.8e-3 at 640x480 (the largest errors sit on the strongest corners, p ~ 0.8; the keypoint rule below uses the error measured AROUND each differing keypoint, not this bound)
HEATMAP_TOL = 8e-3
DESC_TOL = 1e-3     # north_star: descriptors within 1e-3 of the reference graph


def gpu_assignment_scores(read, kp: int, n0: int, n1: int, pair: int = 0) -> np.ndarray:
    """The CUDA path's log-assignment matrix [n0, n1] of pair `pair`, rebuilt from its device buffers exactly as
    argmax_rows_kernel evaluates it: ((sim - lse_row) + (sim - lse_col)) + (lz0 + lz1), fp32.
    `read(what, shape, dtype)` = LightGlue.debug_read or FramePairPipeline.lightglue_debug_read."""
    sim = read("sim", (pair + 1, kp, kp), np.float32)[pair, :n0, :n1]
    lse = read("lse", (2 * pair + 2, kp), np.float32)[2 * pair:]
    lz = read("lz", (2 * pair + 2, kp), np.float32)[2 * pair:]
    return ((sim - lse[0, :n0, None]) + (sim - lse[1, None, :n1])) + (lz[0, :n0, None] + lz[1, None, :n1])


def check_matches(read, kp, lg_weights, m0, ms0, xy0, d0, xy1, d1, w, h, pair=0, mscore_tol=None):
    """matches0 / mscores0 of the CUDA path (for normalised-on-the-fly pixel keypoints xy*, descriptors d*) against the
    oracle on the same inputs.  Returns a small report dict (agreement rate, measured errors) for printing."""
    from oracle import lightglue as olg

    n0, n1 = len(xy0), len(xy1)
    om0, oms0, inter = olg.match(lg_weights, olg.normalize_keypoints(xy0, w, h), d0, olg.normalize_keypoints(xy1, w, h), d1,
                                 return_intermediates=True)
    assert m0.shape == om0.shape
    S = inter["scores"]
    G = gpu_assignment_scores(read, kp, n0, n1, pair)
    err = olg.competitive_score_error(S, G)
    scale = max(1.0, float(np.abs(inter["sim"]).max()))
    assert err <= SCORE_TOL_REL * scale, f"log-assignment scores off by {err:.4g} (similarity logits up to {scale:.3g})"
    rep = olg.disagreement_report(S, om0, oms0, m0, ms0)
    left = olg.explained_by_score_error(rep, err)
    assert not left, (f"{len(left)} of {rep['disagree']} differing matches0 entries are not near-ties (measured score "
                      f"error {err:.4g}): {left[:3]}")
    both = (m0 == om0) & (om0 >= 0)
    ms_err = float(np.abs(ms0[both] - oms0[both]).max()) if both.any() else 0.0
    # mscores0 = exp(score) <= 1:  |d exp(s)| <= exp(s) * (e^err - 1)
    assert ms_err <= np.expm1(err) * 1.05 + 1e-6, f"mscores0 off by {ms_err:.4g} with a score error of {err:.4g}"
    if mscore_tol is not None:
        assert ms_err <= mscore_tol, f"mscores0 off by {ms_err:.4g} (tolerance {mscore_tol})"
    return dict(n=int(n0), oracle_matches=int((om0 >= 0).sum()), differ=rep["disagree"], score_err=err, logit_scale=scale,
                mscore_err=ms_err, max_margin_log=rep["max_margin_log"], om0=om0, oms0=oms0)


def check_keypoints(raw_gpu, raw_ref, xy_gpu, xy_ref, h, w, K, thr=0.005, rb=4):
    """The CUDA path's keypoint set against the oracle's, decision by decision (see module docstring).
    raw_*: heat maps [H', W'] before NMS; xy_*: pixel keypoints [n, 2] as the interface returns them."""
    from oracle import superpoint as osp

    hs, ws = raw_ref.shape
    sx, sy = np.float32(w) / np.float32(ws), np.float32(h) / np.float32(hs)

    def to_hw(xy):
        xy = np.asarray(xy, np.float32).reshape(-1, 2)
        hw = np.stack([np.rint(xy[:, 1] / sy), np.rint(xy[:, 0] / sx)], 1).astype(np.int64)
        back = np.stack([hw[:, 1].astype(np.float32) * sx, hw[:, 0].astype(np.float32) * sy], 1)
        assert np.array_equal(back, xy), "keypoints are not score-map pixels times the reference's scale"
        return hw

    err = float(np.abs(raw_gpu.astype(np.float64) - raw_ref.astype(np.float64)).max())
    assert err <= HEATMAP_TOL, f"heat map off by {err:.4g}"
    rep = osp.keypoint_disagreements(raw_ref, raw_gpu, to_hw(xy_ref), to_hw(xy_gpu), K, thr, rb)
    assert not rep["unexplained"], (f"{len(rep['unexplained'])} of {rep['differ']} differing keypoints are not near-ties: "
                                    f"{rep['unexplained'][:3]}")
    return dict(differ=rep["differ"], n_ref=rep["n_ref"], n_gpu=rep["n_other"], heatmap_err=err)
