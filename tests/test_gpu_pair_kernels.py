"""The CTA-pair kernels (tcgen05.mma.cta_group::2: conv_pipe.cuh kPair, umma_core.cuh kPair) against the single-CTA
kernels they replace.  The switches SSB_SP_PAIR / SSB_LG_PAIR are read once per process, so each variant runs in its own
interpreter on the same inputs.  A pair MMA accumulates every output element over K in the same order as the single-CTA
MMA, so the two variants must agree bit for bit - keypoints, fp16 descriptor rows, match indices and scores - and a
protocol slip between the two CTAs (a halo read before its last store, a stale accumulator) shows up as a difference."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, SP_WEIGHTS

pytestmark = pytest.mark.gpu

_SCRIPT = r"""
import sys, numpy as np
sys.path.insert(0, {root!r})
from superslam_b200 import frontend as fe
from superslam_b200.lightglue_weights import make_random_weights, save_state_dict
from superslam_b200.synth import synth_pair
out, lgw = sys.argv[1], sys.argv[2]
save_state_dict(make_random_weights(7), lgw)
res = {{}}
for tag, (h, w, K, n) in {{"a": (240, 320, 512, 3), "b": (99, 131, 128, 2)}}.items():   # odd size: ragged tiles, odd tile counts
    pairs = [synth_pair(h, w, 50 + i, 100 + 20 * i) for i in range(n)]
    pipe = fe.FramePairPipeline({spw!r}, lgw, K, w, h, max_pairs=n)
    for rep in range(2):                       # eager run, then the captured graph
        o = pipe.process([im for p in pairs for im in p])
    for k in ("count", "xy", "score", "matches0", "mscores0", "has_depth"):
        res[tag + "_" + k] = np.asarray(o[k])
    sp = fe.SuperPoint({spw!r}, K)
    L, R = sp.extract_stereo(*pairs[0])
    lg = fe.LightGlue(lgw, w, h, max_keypoints=K)
    res[tag + "_desc"] = lg.descriptors_to_host(L.descriptors)
np.savez(out, **res)
"""


def _run(tmp_path, name, env_extra):
    out = str(tmp_path / f"{name}.npz")
    env = dict(os.environ, **env_extra)
    script = _SCRIPT.format(root=ROOT, spw=SP_WEIGHTS)
    r = subprocess.run([sys.executable, "-c", script, out, str(tmp_path / f"{name}.ssbw")], env=env, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    return np.load(out)


def test_pair_kernels_equal_single_cta_kernels_bit_for_bit(tmp_path):
    pair = _run(tmp_path, "pair", {"SSB_SP_PAIR": "1", "SSB_LG_PAIR": "1"})
    single = _run(tmp_path, "single", {"SSB_SP_PAIR": "0", "SSB_LG_PAIR": "0"})
    assert set(pair.files) == set(single.files)
    for k in pair.files:
        a, b = pair[k], single[k]
        assert a.shape == b.shape and a.dtype == b.dtype, k
        if k.endswith("_count"):
            assert np.array_equal(a, b), k
    for tag in ("a", "b"):
        cnt = pair[tag + "_count"]
        assert cnt.min() > 10
        for i, n in enumerate(cnt):            # rows beyond the count are unspecified
            assert np.array_equal(pair[tag + "_xy"][i, :n], single[tag + "_xy"][i, :n]), (tag, i)
            assert np.array_equal(pair[tag + "_score"][i, :n], single[tag + "_score"][i, :n]), (tag, i)
        for p in range(len(cnt) // 2):
            n0 = cnt[2 * p]
            assert np.array_equal(pair[tag + "_matches0"][p, :n0], single[tag + "_matches0"][p, :n0]), (tag, p)
            assert np.array_equal(pair[tag + "_mscores0"][p, :n0], single[tag + "_mscores0"][p, :n0]), (tag, p)
            assert np.array_equal(pair[tag + "_has_depth"][p, :n0], single[tag + "_has_depth"][p, :n0]), (tag, p)
        assert np.array_equal(pair[tag + "_desc"], single[tag + "_desc"]), tag
        assert (pair[tag + "_matches0"] >= 0).sum() > 10
