"""Real-weights readiness.  The reference ships no LightGlue weights: its exporter pulls
`superpoint_lightglue.pth` through the un-vendored cvg/LightGlue package (utils/convert_lightglue_to_onnx.py:4-9,69;
scripts/models/download_weights_lightglue.py:2-9), and there is no network here.  The moment that checkpoint is dropped at

    weights/superpoint_lightglue.pth          (or wherever $SSB_LIGHTGLUE_PTH points)

this test converts it (upstream key spelling self_attn.{i}.* / cross_attn.{i}.*, with or without a `matcher.` prefix) and runs
the a11 parity suite on the TRAINED weights with the north-star tolerances: log-assignment scores, matches0 identical up
to near-ties, mscores0 within 1e-3 - on SuperPoint features of a synthetic pair at C2 size and on a ragged random pair.
Until then it is skipped, and the network values stay "parity unpinned" (DESIGN.md section 4)."""
import os

import numpy as np
import pytest

from conftest import ROOT, SP_WEIGHTS

pytestmark = pytest.mark.gpu

PTH = os.environ.get("SSB_LIGHTGLUE_PTH", os.path.join(ROOT, "weights", "superpoint_lightglue.pth"))


@pytest.mark.skipif(not os.path.exists(PTH), reason=f"no LightGlue checkpoint at {PTH} (none ships with the reference; offline)")
def test_trained_checkpoint_meets_the_north_star(tmp_path):
    import torch

    import parity
    from superslam_b200 import frontend as fe
    from superslam_b200.lightglue_weights import normalise_keys, save_state_dict
    from superslam_b200.synth import synth_pair
    from test_gpu_lightglue import _feat

    sd = torch.load(PTH, map_location="cpu", weights_only=True)
    if isinstance(sd, dict) and "state_dict" in sd:
        sd = sd["state_dict"]
    sd = normalise_keys(sd)
    p = str(tmp_path / "lightglue_trained.ssbw")
    save_state_dict(sd, p)
    h, w, K = 480, 640, 1024
    sp = fe.SuperPoint(SP_WEIGHTS, K)
    lg = fe.LightGlue(p, w, h, max_keypoints=K)
    l, r = synth_pair(h, w, 2468)
    L, R = sp.extract_stereo(l, r)
    d0, d1 = lg.descriptors_to_host(L.descriptors), lg.descriptors_to_host(R.descriptors)
    m = lg.match(L.keypoints, L.descriptors, R.keypoints, R.descriptors)
    rep = parity.check_matches(lg.debug_read, lg.kp, sd, m.matches0, m.mscores0, L.keypoints, d0, R.keypoints, d1, w, h,
                               mscore_tol=1e-3)
    print(f"trained weights, C2 pair: {rep['differ']} of {rep['n']} matches0 differ (near-ties), score error {rep['score_err']:.3g}, "
          f"logit scale {rep['logit_scale']:.3g}, mscores0 error {rep['mscore_err']:.3g}, {rep['oracle_matches']} matches")
    assert rep["oracle_matches"] > 100          # a 12 px shifted copy: the trained matcher must find it
    xy0, f0, xy1, f1 = _feat(300, 517, 5)
    m = lg.match(xy0, f0, xy1, f1)
    parity.check_matches(lg.debug_read, lg.kp, sd, m.matches0, m.mscores0, xy0, f0, xy1, f1, w, h, mscore_tol=1e-3)
    assert np.all(m.matches0 < 517)
