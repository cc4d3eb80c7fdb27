#!/usr/bin/env python
"""Stage-by-stage parity report of the CUDA path against the oracle (run on the GPU box):
    python tests/gpu_diag.py [--size HxW] > gpurun_out/diag.txt
Prints one line per stage; never asserts, so one GPU call localises every defect."""
import argparse
import json
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import lightglue as olg  # noqa: E402
from oracle import superpoint as osp  # noqa: E402
from superslam_b200 import frontend as fe  # noqa: E402
from superslam_b200.lightglue_weights import make_random_weights, save_state_dict  # noqa: E402
from superslam_b200.synth import synth_image  # noqa: E402

SPW = os.path.join(ROOT, "superslam_b200", "weights", "superpoint_v1.ssbw")


def read_x32(lg, kp):
    """fp32 residual stream: tile-transposed on the device ([img][row/128][col][row%128]) -> [2, kp, 256]."""
    raw = lg.debug_read("x32", (2, kp // 128, 256, 128), np.float32)
    return raw.transpose(0, 1, 3, 2).reshape(2, kp, 256)


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max()), float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="120x160")
    ap.add_argument("--k", type=int, default=256)
    args = ap.parse_args()
    h, w = map(int, args.size.split("x"))
    K = args.k
    rep = {}
    w_sp = osp.load_weights(SPW)
    imgs = np.stack([synth_image(h, w, 100, max(8, h * w // 480)), synth_image(h, w, 101, max(8, h * w // 480))])
    imgs[1] = np.roll(imgs[0], -12, axis=1)
    sp = fe.SuperPoint(SPW, K)
    t = time.time()
    L, R = sp.extract_stereo(imgs[0], imgs[1])
    print(f"extract_stereo ok in {time.time()-t:.3f}s: n = {len(L.keypoints)}, {len(R.keypoints)}")
    inter = osp.dense_intermediates(imgs, w_sp, fp16_storage=True)
    shapes = {k: v.shape for k, v in inter.items()}
    for name, key in [("conv1b", "conv1b"), ("conv2a", "conv2a"), ("conv2b", "conv2b"),
                      ("conv3a", "conv3a"), ("conv3b", "conv3b"), ("conv4a", "conv4a"), ("conv4b", "conv4b")]:
        got = sp.debug_read(name, shapes[key], np.float16).astype(np.float32)
        print(f"{name:8s} shape {shapes[key]} maxabs/rel {rel(got, inter[key])}")
    b, hc, wc, _ = shapes["convPa"]
    pd = sp.debug_read("convPaDa", (b, hc, wc, 512), np.float16).astype(np.float32)
    print(f"convPa   maxabs/rel {rel(pd[..., :256], inter['convPa'])}")
    print(f"convDa   maxabs/rel {rel(pd[..., 256:], inter['convDa'])}")
    raw = sp.debug_read("scores", inter["raw"].shape, np.float32)
    print(f"scores   maxabs/rel {rel(raw, inter['raw'])}")
    grid = sp.debug_read("grid", inter["grid"].shape, np.float16).astype(np.float32)
    print(f"grid     maxabs/rel {rel(grid, inter['grid'])}")
    # exact index work on the GPU's own heat map
    for i, F in enumerate((L, R)):
        k = osp.nms_select(raw[i], h, w, K, 0.005, 4)
        same_n = len(k["score"]) == len(F.responses)
        same = same_n and np.array_equal(k["xy"], F.keypoints) and np.array_equal(k["score"], F.responses)
        print(f"select[{i}] n_gpu {len(F.responses)} n_oracle {len(k['score'])} exact {same}")
        g16 = sp.debug_read("grid", inter["grid"].shape, np.float16)[i].transpose(2, 0, 1)
        exp = osp.gather_normalize(np.ascontiguousarray(g16), k["cell"])
        lg_dummy = None
        got = np.zeros((len(F.responses), 256), np.float32)
        import ctypes as C
        from superslam_b200 import _lib
        _lib.check(_lib.load().ssb_desc_to_host_f32(0, C.c_void_p(F.descriptors.data), F.descriptors.count, 256,
                                                    got.ctypes.data_as(C.POINTER(C.c_float))))
        if same_n:
            print(f"gather[{i}] bit-exact {np.array_equal(got.astype(np.float16), exp)} maxabs {np.abs(got - exp.astype(np.float32)).max() if len(got) else 0}")
    # fp32 oracle end to end
    ref = osp.extract(imgs, w_sp, K)
    for i, F in enumerate((L, R)):
        a = set(map(tuple, ref[i]["xy"].tolist()))
        bset = set(map(tuple, F.keypoints.tolist()))
        print(f"e2e[{i}] keypoint overlap vs fp32 oracle {len(a & bset)}/{len(a)}")
    # ---- LightGlue ----
    lgw_path = "/tmp/lg_synth.ssbw"
    sd = make_random_weights(7)
    save_state_dict(sd, lgw_path)
    lg = fe.LightGlue(lgw_path, w, h, max_keypoints=K)
    d0 = lg.descriptors_to_host(L.descriptors)
    d1 = lg.descriptors_to_host(R.descriptors)
    k0 = olg.normalize_keypoints(L.keypoints, w, h)
    k1 = olg.normalize_keypoints(R.keypoints, w, h)
    om0, oms0, inter = olg.match(sd, k0, d0, k1, d1, return_intermediates=True)
    kp = (K + 255) // 256 * 256
    n0, n1 = len(k0), len(k1)
    # first self block, stage by stage
    os.environ["SSB_LG_STOP_AFTER"] = "1"
    lg.match(L.keypoints, L.descriptors, R.keypoints, R.descriptors)
    dbg = olg.self_block_debug(sd, 0, k0, d0)
    cs = lg.debug_read("cos", (2, 32, kp), np.float32)[0].T[:n0]   # stored frequency-major [z][32][kp]
    print(f"lg cos maxabs {np.abs(cs - dbg['cos'][:, ::2]).max()}")
    q = lg.debug_read("q", (8, kp, 64), np.float16).astype(np.float32)[:4, :n0]
    k_ = lg.debug_read("k", (8, kp, 64), np.float16).astype(np.float32)[:4, :n0]
    v_ = lg.debug_read("v", (8, kp, 64), np.float16).astype(np.float32)[:4, :n0]
    print(f"lg q {rel(q, dbg['q'])} k {rel(k_, dbg['k'])} v {rel(v_, dbg['v'])}")
    ctx = lg.debug_read("ctx", (2, kp, 256), np.float16).astype(np.float32)[0, :n0]
    h1 = lg.debug_read("h1", (2, kp, 512), np.float16).astype(np.float32)[0, :n0]
    if os.environ.get("SSB_LG_FOLD_OUT", "1") == "0":   # the message only exists in the unfolded graph
        msg = lg.debug_read("msg", (2, kp, 256), np.float16).astype(np.float32)[0, :n0]
        print(f"lg msg {rel(msg, dbg['msg'])}")
    print(f"lg ctx {rel(ctx, dbg['ctx'])} h1 {rel(h1, dbg['h1'])}")
    x32 = read_x32(lg, kp)
    print(f"lg x after self0 {rel(x32[0, :n0], dbg['x'])}")
    for stop, key in [(2, "cross0"), (3, "self1"), (4, "cross1"), (10, "cross4")]:
        os.environ["SSB_LG_STOP_AFTER"] = str(stop)
        lg.match(L.keypoints, L.descriptors, R.keypoints, R.descriptors)
        x32 = read_x32(lg, kp)
        print(f"lg x after {key} img0 {rel(x32[0, :n0], inter[key][0])} img1 {rel(x32[1, :n1], inter[key][1])}")
    del os.environ["SSB_LG_STOP_AFTER"]
    t = time.time()
    m = lg.match(L.keypoints, L.descriptors, R.keypoints, R.descriptors)
    print(f"lg.match ok in {time.time()-t:.3f}s, matches {len(m.query)}")
    x32 = read_x32(lg, kp)
    print(f"lg x32 final img0 maxabs/rel {rel(x32[0, :n0], inter['cross8'][0])} img1 {rel(x32[1, :n1], inter['cross8'][1])}")
    sim = lg.debug_read("sim", (kp, kp), np.float32)
    print(f"lg matches0 equal {np.array_equal(m.matches0, om0)} ({(m.matches0 != om0).sum()} differ of {n0}); "
          f"mscores maxabs {np.abs(m.mscores0 - oms0).max() if n0 else 0}; oracle matches {(om0 >= 0).sum()}")
    print("sim sample", sim[:2, :4], "oracle scores sample", inter["scores"][:2, :4])
    # every disagreement with the fp32 oracle, with the margin of the oracle's decision (a legitimate fp16 flip has a
    # margin within the score error; see oracle/lightglue.py::disagreement_report)
    rep = olg.disagreement_report(inter["scores"], om0, oms0, m.matches0, m.mscores0)
    both = (m.matches0 == om0) & (om0 >= 0)
    err = float(np.abs(np.log(np.maximum(m.mscores0[both], 1e-30)) - np.log(np.maximum(oms0[both], 1e-30))).max()) if both.any() else 0.0
    print(f"lg disagreements {rep['disagree']} of {rep['n']}; largest margin {rep['max_margin']:.4g} "
          f"(log-score error on the agreeing matches {err:.4g})")
    for r in rep["rows"]:
        print(f"   i {r['i']:4d} oracle {r['oracle']:4d} gpu {r['other']:4d} {r['kind']:8s} row_gap {r['row_gap']:.4g} "
              f"col_gap {r['col_gap']:.4g} thr_gap {r['thr_gap']:.4g}")
    # the flip rate against the fixed-precision restatement (fp16 where the CUDA path stores fp16)
    fm0, fms0, finter = olg.match(sd, k0, d0, k1, d1, return_intermediates=True, fp16_storage=True)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import parity
    G = parity.gpu_assignment_scores(lg.debug_read, kp, n0, n1)
    print(f"lg vs fp16-storage oracle: matches0 differ {(m.matches0 != fm0).sum()} of {n0} (fp16 oracle vs fp32 oracle: "
          f"{(fm0 != om0).sum()}); score error GPU-fp16oracle {olg.competitive_score_error(finter['scores'], G):.4g}, "
          f"GPU-fp32oracle {olg.competitive_score_error(inter['scores'], G):.4g}, fp16oracle-fp32oracle "
          f"{olg.competitive_score_error(inter['scores'], finter['scores']):.4g}")


if __name__ == "__main__":
    try:
        main()
    except Exception:
        traceback.print_exc()
        sys.exit(1)
