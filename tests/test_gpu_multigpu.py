"""ssb_mg_* (superslam_b200/csrc/multigpu.cpp) on the real library: the multi-device driver must give, for every pair,
exactly what a single front end gives for that pair - whichever device served it.  With one visible device the same
check runs with that device listed twice (two front ends, two host threads, one GPU)."""
import numpy as np
import pytest

from conftest import SP_WEIGHTS

pytestmark = pytest.mark.gpu


def test_multi_device_driver_equals_single_front_end(tmp_path, lg_weights):
    from superslam_b200 import _lib
    from superslam_b200 import frontend as fe
    from superslam_b200.lightglue_weights import save_state_dict
    from superslam_b200.synth import synth_pair

    lgw = str(tmp_path / "lg.ssbw")
    save_state_dict(lg_weights, lgw)
    n_dev = _lib.load().ssb_device_count()
    assert n_dev >= 1
    devices = list(range(min(n_dev, 8))) if n_dev > 1 else [0, 0]
    h, w, K, pairs = 240, 320, 512, 7
    images = [im for i in range(pairs) for im in synth_pair(h, w, 900 + i, 60 + 10 * i)]
    mg = fe.MultiGpuFrontEnd(SP_WEIGHTS, lgw, K, w, h, devices, max_pairs_per_device=2)
    got = mg.process(images)
    assert [mg.device_of_pair(p) for p in range(pairs)] == [devices[p % len(devices)] for p in range(pairs)]
    single = fe.FramePairPipeline(SP_WEIGHTS, lgw, K, w, h, max_pairs=pairs)
    exp = single.process(images)
    for k in ("count", "xy", "score", "matches0", "mscores0", "has_depth"):
        assert np.array_equal(got[k], exp[k]), k
    assert np.array_equal(np.isnan(got["stereo_ur"]), np.isnan(exp["stereo_ur"]))
    assert np.array_equal(np.nan_to_num(got["stereo_ur"]), np.nan_to_num(exp["stereo_ur"]))
    assert got["has_depth"].sum() > 50
    again = mg.process(images[:6])           # fewer pairs than before, not a multiple of the step
    for k in ("count", "matches0"):
        assert np.array_equal(again[k], exp[k][: (6 if k == "count" else 3)])
    mg.close()
