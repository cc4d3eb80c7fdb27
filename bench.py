#!/usr/bin/env python
"""Benchmark of the hot path: stereo frame-pairs/sec (SuperPoint x2 + LightGlue, 1024 keypoints,
640x480) on N B200s of one node.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 3 --warmup 1     # CPU restatement of the reference path

A "step" is one pass of the whole hot path (conv trunk, heads, NMS, top-K, gather, 9 LightGlue layers,
assignment, stereo post-filter) over one batch of `--pairs` synthetic stereo pairs per GPU.
  value   whole-job pairs/s with the input images already resident in HBM (CUDA events on the
          pipeline's own stream, per step, L2 flushed between steps, max over ranks)
  e2e     the same metric through the public call (FramePairPipeline.process: host u8 images in,
          host keypoints / matches out, H2D + D2H inside the timed region)
  roofline  the dominant tcgen05 kernel, timed live with CUDA events during the timed steps
  cpu_baseline  the oracle (reference restated in fp32 torch) on this box's host cores, bounded sample
Pairs shard across ranks with no data-path collective (weak scaling); rank 0 gathers per-rank
match counts with one NCCL all_gather after the timed region.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SPW = os.path.join(ROOT, "superslam_b200", "weights", "superpoint_v1.ssbw")
CPU_ARM = {
    "reference": "StereoFrontEnd::process over the reference's own SuperPoint / LightGlue C++ classes (compiled in place, "
                 "oracle/_ref/libref_e2e.so); their two TensorRT engines have no CPU form and are served by the fp32 torch "
                 "graphs of the oracle",
    "port": "oracle = the reference's torch graph + restated host logic (the reference itself has no CPU path, TensorRT only)",
}
LG_NOTE = ("LightGlue 9 layers (seeded synthetic LightGlue weights: none ship with the reference); SuperPoint weights = "
           "reference checkpoint")
# BASELINE.json `configs`, in order.  C2 is the configuration the metric is quoted on and the default; the others are
# measured with `--config`.  `pairs` = stereo pairs per GPU and step (C1: 2 * pairs mono frames per step).
CONFIGS = {
    "C1": dict(h=480, w=640, K=1024, pairs=64, extract_only=True,
               metric="frames/sec (SuperPoint only, 1024 kpts, 640x480)", unit="frames/s",
               workload="C1: SuperPoint only, single 640x480 gray frames, K=1024 (tests/test_superpoint_only.cc); "
                        "SuperPoint weights = reference checkpoint"),
    "C2": dict(h=480, w=640, K=1024, pairs=64, metric="stereo frame-pairs/sec (SPx2+LG, 1024 kpts, 640x480)",
               unit="pairs/s", workload="C2: stereo pairs 640x480, K=1024, " + LG_NOTE),
    "C3": dict(h=376, w=1241, K=2048, pairs=32, metric="stereo frame-pairs/sec (SPx2+LG, 2048 kpts, 1241x376)",
               unit="pairs/s", workload="C3: KITTI-size stereo stream 1241x376, K=2048, " + LG_NOTE),
    "C4": dict(h=480, w=752, K=1024, pairs=64, rectify=True, eigenplaces_every=5,
               metric="stereo frame-pairs/sec (rectify + SPx2+LG, 1024 kpts, 752x480, EigenPlaces per keyframe)",
               unit="pairs/s",
               workload="C4: EuRoC-size raw stereo pairs 752x480 rectified on the device, K=1024, EigenPlaces global "
                        "descriptor (512x512 network input, seeded synthetic weights) for every 5th pair's left image, "
                        + LG_NOTE),
    "C5": dict(h=720, w=1280, K=4096, pairs=8, sweep=[1, 2, 4, 8], dynamic=True,
               metric="stereo frame-pairs/sec (SPx2+LG, <=4096 kpts dynamic, 1280x720)", unit="pairs/s",
               workload="C5: synthetic 1280x720 stereo pairs of varying texture density (dynamic keypoint counts up "
                        "to K=4096), micro-batch sweep, " + LG_NOTE),
}
H, W, K = 480, 640, 1024           # set from the chosen config in main()
METRIC = CONFIGS["C2"]["metric"]
WORKLOAD = CONFIGS["C2"]["workload"]


def use_config(name: str):
    global H, W, K, METRIC, WORKLOAD
    c = CONFIGS[name]
    H, W, K, METRIC, WORKLOAD = c["h"], c["w"], c["K"], c["metric"], c["workload"]
    return c


# ---- algorithmic work: 2*MAC counts of the dense contractions, as functions of the configuration
def sp_layer_gflop(h: int, w: int) -> dict:
    """Per image.  640x480: conv1a+1b 23.0, conv2a/2b 5.66, conv3a 2.83, conv3b 5.66, conv4a/4b 1.42, convPa|Da 5.66,
    convPb 0.16, convDb 0.63 = 52.1 GF."""
    h2, w2, h4, w4, hc, wc = h // 2, w // 2, h // 4, w // 4, h // 8, w // 8
    g = lambda px, cin, cout, taps=9: 2.0 * px * taps * cin * cout / 1e9
    return {"sp.conv1ab": g(h * w, 1, 64) + g(h * w, 64, 64), "sp.conv2a": g(h2 * w2, 64, 64), "sp.conv2b": g(h2 * w2, 64, 64),
            "sp.conv3a": g(h4 * w4, 64, 128), "sp.conv3b": g(h4 * w4, 128, 128), "sp.conv4a": g(hc * wc, 128, 128),
            "sp.conv4b": g(hc * wc, 128, 128), "sp.convPaDa": g(hc * wc, 128, 512), "sp.convPb": g(hc * wc, 256, 65, 1),
            "sp.convDb": g(hc * wc, 256, 256, 1)}


def lg_launch_gflop(n0, n1) -> dict:
    """Per launch (one of the 9 layers) and per PAIR with n0 / n1 keypoints; arrays give the sum over pairs.
    Cross attention counts sim once plus two P*V products."""
    n0, n1 = np.asarray(n0, np.float64), np.asarray(n1, np.float64)
    t = n0 + n1
    return {"lg.qkv": float((2 * t * 256 * 768).sum()) / 1e9, "lg.out_proj": float((2 * t * 256 * 256).sum()) / 1e9,
            "lg.ffn1": float((2 * t * 512 * 512).sum()) / 1e9, "lg.ffn2": float((2 * t * 512 * 256).sum()) / 1e9,
            "lg.ffn": float((2 * t * (512 * 512 + 512 * 256)).sum()) / 1e9,
            "lg.qkv_cross": float((2 * t * 256 * 512).sum()) / 1e9, "lg.to_out": float((2 * t * 256 * 256).sum()) / 1e9,
            "lg.attn_self": float((4 * 256 * (n0 * n0 + n1 * n1)).sum()) / 1e9,
            "lg.attn_cross": float((6 * 256 * n0 * n1).sum()) / 1e9,
            "lg.final_proj": float((2 * t * 256 * 256).sum()) / 1e9, "lg.sim": float((2 * 256 * n0 * n1).sum()) / 1e9}


def pair_gflop(h: int, w: int, n0, n1) -> float:
    """Algorithmic GFLOP of SP x2 + LG summed over the pairs with keypoint counts n0 / n1 (arrays).
    640x480 at 1024/1024: 104.2 + 80.5 = 184.7 per pair."""
    g = lg_launch_gflop(n0, n1)
    per_layer = (g["lg.qkv"] + g["lg.out_proj"] + g["lg.ffn"] + g["lg.qkv_cross"] + g["lg.to_out"] + g["lg.ffn"] +
                 g["lg.attn_self"] + g["lg.attn_cross"])
    return 2 * sum(sp_layer_gflop(h, w).values()) * np.size(n0) + 9 * per_layer + g["lg.final_proj"] + g["lg.sim"]


_JSON_FD = None


def claim_stdout() -> None:
    """stdout carries exactly one JSON line: everything libraries print to fd 1 (e.g. the NCCL version banner at
    communicator creation) is sent to stderr instead."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def lg_weights_path(rank: int) -> str:
    from superslam_b200.lightglue_weights import make_random_weights, save_state_dict

    p = f"/tmp/ssb_bench_lightglue_r{rank}.ssbw"
    save_state_dict(make_random_weights(7), p)
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_pair_seconds(n_iters: int, warmup: int, budget_s: float = 25.0, reference_classes: bool = False,
                        extract_only: bool = False):
    """One stereo pair through the CPU oracle (fp32 torch restatement of the reference graph +
    restated host logic).  The intra-op thread count is calibrated first (8, 16, ... up to every host
    thread; the fastest wins: on a 128-thread box the oracle's many small ops run several times slower with
    128 threads than with 16), then up to `n_iters` pairs are timed, stopping early once `budget_s` seconds
    of timed work have passed.  Returns (median seconds per pair, threads used, pairs timed, kind, host cores).
    `extract_only` (config C1): the pair call stops after SuperPoint (two frames per call)."""
    import torch

    from oracle import frontend as ofe
    from oracle import lightglue as olg
    from oracle import superpoint as osp
    from superslam_b200.lightglue_weights import make_random_weights
    from superslam_b200.synth import synth_pair

    wsp = osp.load_weights(SPW)
    wlg = make_random_weights(7)
    l, r = synth_pair(H, W, 1234)

    def one_pair_port():
        t = time.perf_counter()
        res = osp.extract(np.stack([l, r]), wsp, K)
        if extract_only:
            return time.perf_counter() - t
        m0, ms0 = olg.match(wlg, olg.normalize_keypoints(res[0]["xy"], W, H), res[0]["desc"],
                            olg.normalize_keypoints(res[1]["xy"], W, H), res[1]["desc"])
        q, tr, _ = ofe.dmatches(m0, ms0)
        ofe.stereo_postfilter(res[0]["xy"], res[1]["xy"], q, tr)
        return time.perf_counter() - t

    one_pair, kind = one_pair_port, "port"
    if reference_classes and not extract_only:   # --impl reference: the reference's own wrapper classes when they were compiled in place
        try:                # (oracle/_ref/libref_e2e.so, see oracle/Makefile); the GPU arm's cpu_baseline keeps the plain port
            one_pair_ref = _reference_classes_pair(osp, olg, wsp, wlg, l, r)
            if one_pair_ref is not None:
                one_pair, kind = one_pair_ref, "reference"
        except Exception as e:  # the baseline must not depend on it
            sys.stderr.write(f"bench: reference classes unavailable ({e}); timing the oracle port\n")

    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    best_t, best_c = None, None
    for c in sorted({min(ncpu, x) for x in (8, 16, 32, 64, ncpu)}):
        torch.set_num_threads(c)
        if best_t is None:
            for _ in range(max(1, warmup)):
                one_pair()  # first-touch / allocator warm-up
        t = one_pair()
        if best_t is None or t < best_t:
            best_t, best_c = t, c
        elif t > 1.5 * best_t:
            break
        if best_t > 15.0:
            break  # a box this slow: keep the bounded sample bounded, do not try further counts
    torch.set_num_threads(best_c)
    times, spent = [], 0.0
    while len(times) < max(1, n_iters) and (not times or spent < budget_s):
        times.append(one_pair())
        spent += times[-1]
    return float(np.median(times)), best_c, len(times), kind, ncpu


_E2E_KEEP = []   # ctypes callbacks and buffers of _reference_classes_pair must outlive the call


def _reference_classes_pair(osp, olg, wsp, wlg, l, r):
    """StereoFrontEnd::process over the reference's own SuperPoint / LightGlue classes (src/*.cc compiled in place into
    oracle/_ref/libref_e2e.so), their two TensorRT engines - which have no CPU form - served by the fp32 torch graphs of the
    oracle, CUDA runtime calls on host memory.  Returns a callable that times one pair, or None when the library is absent."""
    import ctypes as C

    path = os.path.join(ROOT, "oracle", "_ref", "libref_e2e.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    fp, ip, u16p = C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_uint16)
    SP = C.CFUNCTYPE(None, fp, C.c_int, C.c_int, C.c_int, fp, u16p)
    LG = C.CFUNCTYPE(None, fp, C.c_int, u16p, fp, C.c_int, u16p, ip, fp)
    GA = C.CFUNCTYPE(None, u16p, C.c_int, C.c_int, C.c_int, ip, ip, C.c_int, u16p)

    def arr(ptr, shape, dtype):
        return np.ctypeslib.as_array(ptr, shape=(int(np.prod(shape)),)).view(dtype).reshape(shape)

    @SP
    def sp_infer(image, b, h, w, scores, desc):
        s, grid, _ = osp.dense_forward(arr(image, (b, 1, h, w), np.float32).copy(), wsp, fp16_storage=False)
        arr(scores, s.shape, np.float32)[:] = s
        arr(desc, grid.shape, np.uint16)[:] = grid.astype(np.float16).view(np.uint16)

    @LG
    def lg_infer(k0, n0, d0, k1, n1, d1, m0, ms0):
        a = olg.match(wlg, arr(k0, (n0, 2), np.float32).copy(), arr(d0, (n0, 256), np.uint16).view(np.float16).copy(),
                      arr(k1, (n1, 2), np.float32).copy(), arr(d1, (n1, 256), np.uint16).view(np.float16).copy())
        arr(m0, (n0,), np.int32)[:] = a[0]
        arr(ms0, (n0,), np.float32)[:] = a[1]

    @GA
    def gather(grid, c, gh, gw, cell_h, cell_w, n, out):
        cell = np.stack([arr(cell_h, (n,), np.int32), arr(cell_w, (n,), np.int32)], 1)
        arr(out, (n, c), np.uint16)[:] = osp.gather_normalize(arr(grid, (c, gh, gw), np.uint16).view(np.float16).copy(),
                                                              cell).view(np.uint16)

    lib.ref_e2e_set_hooks(sp_infer, lg_infer, gather)
    lib.ref_e2e_create.restype = C.c_void_p
    lib.ref_e2e_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_float, ip]
    lib.ref_e2e_process_only.restype = C.c_int
    lib.ref_e2e_process_only.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, ip]
    engines = []
    for tag in (b"superpoint", b"lightglue"):
        p = f"/tmp/ssb_bench_{tag.decode()}_{os.getpid()}.engine"
        with open(p, "wb") as f:
            f.write(tag + b" stand-in engine")
        engines.append(p.encode())
    st = C.c_int(-1)
    h = lib.ref_e2e_create(engines[0], engines[1], K, 0.005, 4, W, H, 1.0, C.byref(st))
    if st.value != 3:
        return None
    left, right = np.ascontiguousarray(l), np.ascontiguousarray(r)
    _E2E_KEEP.extend([lib, sp_infer, lg_infer, gather, left, right])

    def one_pair():
        nd = C.c_int(0)
        t = time.perf_counter()
        n = lib.ref_e2e_process_only(h, left.ctypes.data, right.ctypes.data, H, W, left.strides[0], C.byref(nd))
        dt = time.perf_counter() - t
        if n <= 0:
            raise RuntimeError("the reference classes returned no keypoints")
        return dt

    one_pair()   # proves the path before it is chosen
    return one_pair


def bench_latency(device: int, rank: int, iters: int = 60, warm: int = 8):
    """single-pair latency in the reference's natural mode - one synchronous
    IFeatureExtractor::extract_stereo + IFeatureMatcher::match per frame (src/StereoFrontEnd.cc:14,33), host
    images in, host keypoints / matches out - with the reference's own profile labels
    (fe_extract_stereo, fe_lg_stereo_match), plus the same pair through the one-call pipeline (pairs = 1)."""
    from superslam_b200 import frontend as fe
    from superslam_b200.synth import synth_pair

    sp = fe.SuperPoint(SPW, K, device=device)
    lg = fe.LightGlue(lg_weights_path(rank), W, H, max_keypoints=K, device=device)
    pipe1 = fe.FramePairPipeline(SPW, lg_weights_path(rank), K, W, H, max_pairs=1, device=device)
    pairs = [synth_pair(H, W, 9000 + i) for i in range(4)]
    t_ext, t_match, t_pipe = [], [], []
    for it in range(warm + iters):
        l, r = pairs[it % len(pairs)]
        t0 = time.perf_counter()
        L, R = sp.extract_stereo(l, r)
        t1 = time.perf_counter()
        lg.match(L.keypoints, L.descriptors, R.keypoints, R.descriptors)
        t2 = time.perf_counter()
        pipe1.process([l, r])
        t3 = time.perf_counter()
        if it >= warm:
            t_ext.append(t1 - t0), t_match.append(t2 - t1), t_pipe.append(t3 - t2)
        del L, R

    def q(v):
        v = np.asarray(v) * 1e3
        return {"p50_ms": round(float(np.percentile(v, 50)), 3), "p95_ms": round(float(np.percentile(v, 95)), 3)}

    both = np.asarray(t_ext) + np.asarray(t_match)
    return {"workload": "1 pair 640x480, K=1024 per call, host in / host out, synchronous", "iterations": iters,
            "fe_extract_stereo": q(t_ext), "fe_lg_stereo_match": q(t_match), "extract_plus_match": q(both),
            "pipeline_process_1_pair": q(t_pipe)}


def bench_eigenplaces(lib, device: int, steps: int = 10, batch: int = 8):
    """the EigenPlaces global descriptor (ResNet18 + GeM + FC on the tcgen05 conv
    core) for `batch` 752x480 keyframes per call, host images in, host descriptors out (the public call)."""
    from superslam_b200 import frontend as fe
    from superslam_b200.eigenplaces_weights import make_random_weights, save_state_dict
    from superslam_b200.synth import synth_pair

    path = f"/tmp/ssb_bench_eigenplaces_d{device}.ssbw"
    save_state_dict(make_random_weights(11), path)
    ep = fe.EigenPlaces(path, 512, 512, max_batch=batch, device=device)
    imgs = [synth_pair(480, 752, 4000 + i)[0] for i in range(batch)]
    for _ in range(3):
        d = ep.compute_global_descriptors(imgs)
    t0 = time.perf_counter()
    for _ in range(steps):
        d = ep.compute_global_descriptors(imgs)
    dt = (time.perf_counter() - t0) / steps
    lib.ssb_profile_enable(1)
    for _ in range(3):
        ep.compute_global_descriptors(imgs)
        lib.ssb_profile_collect()
    lib.ssb_profile_enable(0)
    buf = C.create_string_buffer(1 << 16)
    lib.ssb_profile_report(buf, len(buf))
    k = {}
    for ln in buf.value.decode().splitlines():
        name, cnt, ms = ln.split()
        if name.startswith("ep."):
            k[name] = round(float(ms) / 3, 4)
    gf = 18.95   # 2*MAC of ResNet18 up to layer4 at 512x512 (stem 1.23, layer1 4.83, layers 2-4 4.30 each)
    dev_ms = sum(k.values())
    return {"images_per_s": batch / dt, "batch": batch, "workload": "C4 keyframes 752x480 gray -> 512x512 network input",
            "ms_per_call_e2e": dt * 1e3, "device_ms_per_call": dev_ms, "kernel_ms_per_call": k,
            "algorithmic_gflop_per_image": gf, "tflops_device": gf * batch / dev_ms if dev_ms else None,
            "descriptor_norm": float(np.linalg.norm(d[0])) if len(d) else None}


def run_reference(args, rank: int, cfg_name: str):
    if rank != 0:
        return
    cfg = CONFIGS[cfg_name]
    sec, threads, timed, kind, ncpu = oracle_pair_seconds(max(1, args.steps), min(1, args.warmup), budget_s=90.0,
                                                          reference_classes=True,
                                                          extract_only=bool(cfg.get("extract_only")))
    per_step = 2 if cfg.get("extract_only") else 1    # C1 counts frames: one timed step = the two frames of a pair call
    v = per_step / sec
    line = {
        "metric": METRIC, "value": v, "unit": cfg["unit"], "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": {"workload": WORKLOAD, "name": cfg_name, "pairs_per_step": 1,
                   "sample": "each step = one whole pair of that workload on the host cores (a bounded sample of the "
                             "device step)"},
        "cpu_baseline": {"value": v, "unit": cfg["unit"], "cores": threads, "host_cores": ncpu, "kind": kind,
                         "sample": f"{timed} pair(s) timed after thread-count calibration and warm-up; " + CPU_ARM[kind]},
        "e2e": {"value": v, "unit": cfg["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def identity_rectify_maps(h: int, w: int, seed: int):
    """Mild synthetic rectification maps (identity plus a smooth sub-pixel warp): the remap kernel does the full
    bilinear work, the image content stays that of the synthetic pair."""
    yy, xx = np.meshgrid(np.arange(h, dtype=np.float32), np.arange(w, dtype=np.float32), indexing="ij")
    rng = np.random.default_rng(seed)
    a, b = rng.uniform(0.2, 0.6, 2).astype(np.float32)
    mx = xx + a * np.sin(yy / np.float32(37.0)).astype(np.float32)
    my = yy + b * np.cos(xx / np.float32(53.0)).astype(np.float32)
    return mx.astype(np.float32), my.astype(np.float32)


def make_images(cfg, P: int, rank: int, world: int, step: int = 0):
    """The 2P images of one step of rank `rank`: slot s holds pair (rank + world * (step * P + s)) of the round-robin
    stream (superslam_b200/sharding.py).  C5 draws each pair's texture density at random: dynamic keypoint counts."""
    from superslam_b200.synth import default_n_shapes, synth_pair

    images, index = [], []
    for s_ in range(P):
        g = rank + world * (step * P + s_)
        n_shapes = None
        if cfg.get("dynamic"):
            frac = np.random.default_rng(777 + g).uniform(0.15, 1.0)
            n_shapes = max(8, int(default_n_shapes(H, W) * frac))
        l, r = synth_pair(H, W, 1234 + g, n_shapes)
        images += [l, r]
        index.append(g)
    return images, index


def timed_steps(pipe, dev_images, P, steps, flush, torch, extra=None):
    """`steps` steps with the images resident in HBM: CUDA events on the pipeline stream around each step (one CUDA-graph
    replay), L2 flushed before each.  `extra()` (optional) is host-timed work that belongs to the step (C4: EigenPlaces)."""
    ms, extra_ms = [], []
    for _ in range(steps):
        flush.zero_()
        torch.cuda.synchronize()
        pipe.event_record(0)
        pipe.enqueue_device(dev_images, P, H, W)
        pipe.event_record(1)
        pipe.sync()
        ms.append(pipe.event_elapsed_ms(0, 1))
        if extra is not None:
            t0 = time.perf_counter()
            extra()
            extra_ms.append((time.perf_counter() - t0) * 1e3)
    return ms, extra_ms


def spread(values):
    v = np.asarray(values, np.float64)
    return {"min": float(v.min()), "median": float(np.median(v)), "max": float(v.max()), "n": int(v.size)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS), help="BASELINE.json configuration (default: the "
                    "one the metric is quoted on)")
    ap.add_argument("--pairs", type=int, default=0, help="stereo pairs per GPU per step (default: the config's)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the latency / EigenPlaces / live-pipeline sections")
    args = ap.parse_args()
    claim_stdout()
    cfg = use_config(args.config)

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if args.steps > 5:
            args.steps = 5  # bounded sample of the workload: whole pairs, at most 5 (and at most ~90 s)
        run_reference(args, rank, args.config)
        return

    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from superslam_b200 import _lib
    from superslam_b200 import frontend as fe
    from superslam_b200 import sharding

    lib = _lib.load()
    P = args.pairs or cfg["pairs"]
    units_per_pair = 2 if cfg.get("extract_only") else 1          # C1 counts frames
    pipe = fe.FramePairPipeline(SPW, lg_weights_path(rank), K, W, H, max_pairs=P, device=local)
    if cfg.get("extract_only"):
        pipe.set_extract_only(True)
    rect = None
    if cfg.get("rectify"):   # EuRoC flow: raw images in, rectified on the device in front of SuperPoint
        rect = [fe.Rectifier(*identity_rectify_maps(H, W, 40 + i), (H, W), max_images=P, device=local) for i in range(2)]
        pipe.set_rectifiers(rect[0], rect[1])
    images, pair_index = make_images(cfg, P, rank, world)
    dev_images = pipe.upload(images)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")  # > 126 MB L2
    ep, ep_imgs = None, None
    if cfg.get("eigenplaces_every"):
        from superslam_b200.eigenplaces_weights import make_random_weights as ep_weights, save_state_dict as ep_save

        ep_path = f"/tmp/ssb_bench_eigenplaces_r{rank}.ssbw"
        ep_save(ep_weights(11), ep_path)
        ep_imgs = [images[2 * s_] for s_ in range(0, P, cfg["eigenplaces_every"])]
        ep = fe.EigenPlaces(ep_path, 512, 512, max_batch=len(ep_imgs), device=local)
    extra = (lambda: ep.compute_global_descriptors(ep_imgs)) if ep is not None else None

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ----
    W_ = max(3, args.warmup)
    timed_steps(pipe, dev_images, P, W_, flush, torch, extra)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    launches0 = fe.kernel_launch_count()
    t_wall = time.perf_counter()
    step_ms, extra_ms = timed_steps(pipe, dev_images, P, args.steps, flush, torch, extra)
    barrier()
    wall = time.perf_counter() - t_wall
    launches = fe.kernel_launch_count() - launches0
    # per-kernel CUDA-event timing for the roofline: the same steps once more, launched eagerly with an
    # event after every kernel on the pipeline stream (a replayed CUDA graph cannot carry timing events)
    lib.ssb_profile_enable(1)
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        pipe.enqueue_device(dev_images, P, H, W)
        pipe.sync()
        lib.ssb_profile_collect()
    lib.ssb_profile_enable(0)
    clocks = sampler.stop()
    buf = C.create_string_buffer(1 << 16)
    lib.ssb_profile_report(buf, len(buf))
    prof = {}
    for ln in buf.value.decode().splitlines():
        name, cnt, ms = ln.split()
        prof[name] = (int(cnt), float(ms))
    out = pipe.fetch(P)

    dev_total_ms = float(sum(step_ms) + sum(extra_ms))
    t = torch.tensor([dev_total_ms], dtype=torch.float64, device=f"cuda:{local}")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_ms = float(t.item())
    value = world * P * units_per_pair * args.steps / (max_ms / 1e3)
    per_step = [P * units_per_pair / ((a + (extra_ms[i] if extra_ms else 0.0)) / 1e3) for i, a in enumerate(step_ms)]

    # ---- end to end through the public call, host buffers in / out ----
    # Streaming API (ssb_fe_submit / ssb_fe_collect): every step uploads its 2P images from pinned host memory
    # and reads its results back to the host; the upload of step i+1 is in flight while step i computes.  With
    # N > 1 ranks every collected step is followed by the result gather of SURVEY 8e INSIDE the timed loop: the step's
    # padded records (pair index, counts, matches0, mscores0, has_depth) go through one NCCL all_gather and land on
    # the host of every rank.
    pinned = [torch.from_numpy(im).pin_memory() for im in images]
    images = [t_.numpy() for t_ in pinned]
    gathered_pairs, record_bytes = 0, 0
    gather_dev = None

    def gather(res, step):
        nonlocal gathered_pairs, record_bytes, gather_dev
        if dist is None or cfg.get("extract_only"):
            return None
        rec = sharding.pack_records_block(rank + world * step * P, world, res["count"], res["matches0"], res["mscores0"],
                                          res["has_depth"], K)
        tr = torch.from_numpy(rec).to(f"cuda:{local}", non_blocking=False)
        if gather_dev is None or gather_dev.shape[1:] != tr.shape:
            gather_dev = torch.empty((world,) + tuple(tr.shape), dtype=tr.dtype, device=tr.device)
        dist.all_gather_into_tensor(gather_dev, tr)
        allrec = gather_dev.cpu().numpy()       # every rank holds every rank's records on the host
        gathered_pairs = int((allrec[:, :, 0] >= 0).sum())
        record_bytes = int(rec.nbytes)
        return allrec

    def e2e_region(steps):
        """Exactly `steps` steps submitted AND collected inside the timed region (pipeline fill and drain included):
        the upload of step i+1 is in flight while step i computes and step i-1's results are gathered."""
        t0 = time.perf_counter()
        pipe.submit(images)
        res_ = None
        for i in range(steps):
            if i + 1 < steps:
                pipe.submit(images)
            res_ = pipe.collect()
            if extra is not None:
                extra()
            gather(res_, i)
        torch.cuda.synchronize()
        return time.perf_counter() - t0, res_

    for _ in range(2):
        pipe.process(images)
    e2e_region(7)                # both image buffers seen three times: eager, capture, replay
    barrier()
    e2e_s, res = e2e_region(args.steps)
    last_records = gather(res, args.steps - 1)
    e2e_repeats = [e2e_s]
    for _ in range(2):
        barrier()
        e2e_repeats.append(e2e_region(args.steps)[0])
    # the same through the synchronous call (upload -> kernels -> read-back, nothing overlapped)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pipe.process(images)
        if extra is not None:
            extra()
    e2e_sync_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s, e2e_sync_s] + e2e_repeats, dtype=torch.float64, device=f"cuda:{local}")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    units = world * P * units_per_pair * args.steps
    e2e_value = units / float(t[0].item())
    e2e_sync_value = units / float(t[1].item())
    e2e_rep_values = [units / float(x) for x in t[2:].tolist()]
    h2d = 2 * P * H * W + (sum(im.nbytes for im in ep_imgs) if ep_imgs else 0)
    d2h = sum(v.nbytes for v in res.values()) + (len(ep_imgs) * 512 * 4 if ep_imgs else 0)

    # per-rank totals + a check of the gathered records of the last step against this rank's own results
    n_left, n_right = out["count"][0::2].astype(np.int64), out["count"][1::2].astype(np.int64)
    matches = torch.tensor([int((out["matches0"] >= 0).sum()), int(out["has_depth"].sum()), int(out["count"].sum())],
                           dtype=torch.int64, device=f"cuda:{local}")
    gathered = [matches]
    if dist is not None:
        gathered = [torch.zeros_like(matches) for _ in range(world)]
        dist.all_gather(gathered, matches)
    records = {"inside_timed_e2e": dist is not None, "pairs_per_gather": gathered_pairs,
               "record_bytes_per_rank_per_step": record_bytes}
    if last_records is not None:
        try:
            got = sharding.unpack_records(last_records, K)
            mine = [rank + world * ((args.steps - 1) * P + s_) for s_ in range(P)]
            ok = all(np.array_equal(got[g]["matches0"], res["matches0"][s_]) for s_, g in enumerate(mine))
            records.update({"pairs_in_last_gather": len(got), "own_pairs_intact": bool(ok),
                            "pair_index": "slot s of rank r at step i = pair r + world * (i * P + s), the same rule "
                                          "make_images() seeds the pair with"})
        except Exception as e:  # never at the cost of the headline line
            records["error"] = str(e)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained"
        peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
        sp_gf = {k: v * 2 * P for k, v in sp_layer_gflop(H, W).items()}           # per launch: 2P images
        lg_gf = lg_launch_gflop(n_left, n_right) if not cfg.get("extract_only") else {}
        gflop = {**sp_gf, **lg_gf}
        total_prof_ms = sum(ms for _, ms in prof.values()) or 1.0
        dom = max(prof.items(), key=lambda kv: kv[1][1])[0] if prof else None
        traffic_file = None
        for cand in ("ncu_traffic_r02.json", "ncu_traffic_r01.json"):
            if os.path.exists(os.path.join(ROOT, "profiles", cand)):
                traffic_file = cand
                break
        roof = None
        if dom is not None:
            cnt, ms = prof[dom]
            avg_ms = ms / max(1, cnt)
            if dom in gflop:
                gf = gflop[dom]
                traffic = None
                try:  # DRAM bytes per launch from the committed ncu capture (per image / per pair at C2), times this run's batch
                    tr = json.load(open(os.path.join(ROOT, "profiles", traffic_file))).get(dom)
                    if tr and args.config == "C2":
                        traffic = int(tr["bytes"] * (2 * P if tr["per"] == "image" else P))
                except Exception:
                    pass
                roof = {"kernel": dom, "bound": "tensor", "achieved": gf / avg_ms, "peak": peak_tf, "unit": "TFLOP/s",
                        "frac": gf / avg_ms / peak_tf, "traffic": traffic,
                        "traffic_source": f"committed ncu capture profiles/{traffic_file} (dram__bytes_read.sum + "
                                          "dram__bytes_write.sum per image/pair at 64 pairs/step of C2, x this run's "
                                          "batch; not measured in this run)",
                        "avg_launch_ms": avg_ms, "share_of_step": ms / total_prof_ms, "peak_source": peak_src,
                        "algorithmic_gflop_per_launch": gf}
                if dom == "sp.conv1ab":   # what keeps this kernel off the tensor roof (profiles/README.md, round 2)
                    roof["limiter"] = ("shared-memory port: an N = 64 tcgen05.mma reads 4 KB of A + 2 KB of B per 32 cycles "
                                       "of math = 48 cycles at 128 B/clk (measured 48.0; CTA pairs 43.0, "
                                       "tools/umma2_probe.cu), plus halo / im2col / staging traffic on the same port")
            else:
                roof = {"kernel": dom, "bound": "hbm", "achieved": None, "peak": peak_gbs,
                        "unit": "GB/s", "frac": None, "traffic": None, "avg_launch_ms": avg_ms,
                        "share_of_step": ms / total_prof_ms, "peak_source": peak_src}
        shares = {k: round(v[1] / total_prof_ms, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:12]}
        tensor_kernels = {}
        for name, (cnt, ms) in prof.items():
            gfl = gflop.get(name, 0)
            if gfl > 0 and ms > 0:
                tensor_kernels[name] = {"avg_launch_ms": round(ms / cnt, 5), "tflops": round(gfl / (ms / cnt), 1),
                                        "frac_of_peak": round(gfl / (ms / cnt) / peak_tf, 4)}
        # whole step against the tensor roofline: algorithmic FLOPs of the step (this config, the step's actual keypoint
        # counts) / the device-timed step
        if cfg.get("extract_only"):
            step_gf = sum(sp_gf.values())
        else:
            step_gf = pair_gflop(H, W, n_left, n_right)
        pipeline_roof = {"algorithmic_gflop_per_step": step_gf, "gflop_per_unit": step_gf / (P * units_per_pair),
                         "tflops": step_gf / (float(np.mean(step_ms))), "peak": peak_tf,
                         "frac": step_gf / float(np.mean(step_ms)) / peak_tf,
                         "attention_frac_of_lightglue_flops": (
                             9 * (lg_gf["lg.attn_self"] + lg_gf["lg.attn_cross"]) /
                             max(1e-9, step_gf - sum(sp_gf.values())) if lg_gf else None),
                         "note": "EigenPlaces / rectification time of C4 is inside `value` but its FLOPs are not counted here"}
        # the HBM-bound pieces: algorithmic bytes per launch / measured launch time against the measured copy
        # bandwidth.  Bytes per unit as stated in DESIGN.md section 3.
        hbm_kernels = {}
        try:
            hc8, wc8 = (H // 8) * 8, (W // 8) * 8
            nm = float((n_left * n_right).sum())
            per_launch_bytes = {"sp.nms": 4 * hc8 * wc8 * 2 * P, "sp.gather": int(out["count"].sum()) * 1024,
                                "lg.lse": 2 * nm * 4, "lg.argmax": 2 * nm * 4, "lg.assign": 2 * nm * 4}
            for name, nbytes in per_launch_bytes.items():
                if name in prof and prof[name][1] > 0:
                    cnt, ms = prof[name]
                    gbs = nbytes / (ms / cnt) / 1e6
                    hbm_kernels[name] = {"avg_launch_ms": round(ms / cnt, 5), "algorithmic_mb_per_launch": round(nbytes / 1e6, 2),
                                         "gb_per_s": round(gbs, 1), "frac_of_peak": round(gbs / peak_gbs, 4)}
        except Exception as e:  # a reporting extra must never cost the headline line
            hbm_kernels = {"error": str(e)}
        extras = world == 1 and not args.no_extras and args.config == "C2"
        eigen = latency = live = None
        if extras:
            try:
                eigen = bench_eigenplaces(lib, local)
            except Exception as e:  # the headline line must not depend on the "next" rows
                eigen = {"error": str(e)}
            try:
                latency = bench_latency(local, rank)
            except Exception as e:
                latency = {"error": str(e)}
            try:
                live = bench_live_pipeline(pipe, dev_images, P, flush, torch, max(5, args.steps // 2))
            except Exception as e:
                live = {"error": str(e)}
        sweep = None
        if world == 1 and cfg.get("sweep") and not args.no_extras:
            try:
                sweep = bench_sweep(cfg, local, rank, flush, torch, max(3, args.steps // 4))
            except Exception as e:
                sweep = {"error": str(e)}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            sec, threads, timed, kind, ncpu = oracle_pair_seconds(5, 1, budget_s=20.0,
                                                                  extract_only=bool(cfg.get("extract_only")))
            cpu = {"value": units_per_pair / sec, "unit": cfg["unit"], "cores": threads, "host_cores": ncpu, "kind": kind,
                   "sample": f"{timed} pair(s) of the same {args.config} workload after warm-up and thread-count "
                             f"calibration ({threads} torch threads of {ncpu} host cores); " + CPU_ARM[kind]}
        line = {
            "metric": METRIC, "value": value, "unit": cfg["unit"], "n_gpus": world, "steps": args.steps,
            "warmup": W_, "ms_per_step": max_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16 (fp32 accumulate)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "name": args.config,
                       "pairs_per_step_per_gpu": P, "l2": "flushed between steps (256 MiB memset)",
                       "keypoints_per_image": {"mean": float(out["count"].mean()), "min": int(out["count"].min()),
                                               "max": int(out["count"].max())},
                       "timing": "CUDA events per step on the pipeline stream (CUDA-graph replay), max over ranks; "
                                 "roofline/kernel shares from an eager re-run of the same steps with an event "
                                 "after every kernel" +
                                 ("; the EigenPlaces call of every step is host-timed after the step's sync and added "
                                  "(serial: conservative)" if ep is not None else "")},
            "value_spread_per_step": spread(per_step),
            "e2e": {"value": e2e_value, "unit": cfg["unit"], "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "FramePairPipeline.submit/collect (streaming: H2D of step i+1 under the kernels of step i)" +
                           ("; + one NCCL all_gather of the step's result records to every rank's host per step"
                            if dist is not None else ""),
                    "repeats": spread(e2e_rep_values),
                    "synchronous_process_value": e2e_sync_value},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "pipeline_roofline": pipeline_roof,
            "kernel_time_shares": shares,
            "tensor_kernels": tensor_kernels,
            "hbm_kernels": hbm_kernels,
            "kernel_ms_per_step": {k: round(v[1] / args.steps, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])},
            "cpu_baseline": cpu,
            "latency_single_pair": latency,
            "live_pipeline": live,
            "eigenplaces": eigen,
            "micro_batch_sweep": sweep,
            "wall_s_timed_region": wall,
            "results": {"per_rank_[matches,has_depth,keypoints]": [g.tolist() for g in gathered],
                        "gathered_records": records},
        }
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


def bench_live_pipeline(pipe, dev_images, P, flush, torch, steps):
    """the live pipeline is SP x2 + LG x2 per frame (stereo match + last-keyframe <-> left tracking match,
    src/VoEstimator.cc:240-246).  Same resident images, tracking enabled, every stream's keyframe = its own previous
    frame (promoted once): pairs/s of the chained graph."""
    pipe.enable_tracking(True)
    try:
        for _ in range(3):
            pipe.enqueue_device(dev_images, P, H, W)
        pipe.fetch(P)
        pipe.promote_keyframes(P)
        for _ in range(3):   # eager, capture, replay with the keyframes in place
            pipe.enqueue_device(dev_images, P, H, W)
        pipe.sync()
        ms, _ = timed_steps(pipe, dev_images, P, steps, flush, torch)
        pipe.fetch(P)
        trk = pipe.tracking_results(P)
        return {"workload": "C2 pairs, tracking chain on: SP x2 + LG (stereo) + LG (keyframe <-> left) per pair, one graph",
                "pairs_per_s": P * len(ms) / (sum(ms) / 1e3), "ms_per_step": float(np.mean(ms)), "steps": len(ms),
                "kernel_launches_per_step": pipe.kernel_launches_per_call(P),
                "tracking_matches_per_pair": float((trk["track_matches0"] >= 0).sum() / P),
                "usable_per_pair": float(trk["track_usable"].sum() / P)}
    finally:
        pipe.enable_tracking(False)


def bench_sweep(cfg, device, rank, flush, torch, steps):
    """C5: micro-batch sweep - pairs per step in cfg['sweep'], device-resident, same dynamic-count images."""
    from superslam_b200 import frontend as fe

    res = {}
    for mb in cfg["sweep"]:
        p = fe.FramePairPipeline(SPW, lg_weights_path(rank), K, W, H, max_pairs=mb, device=device)
        imgs, _ = make_images(cfg, mb, rank, 1)
        dev = p.upload(imgs)
        timed_steps(p, dev, mb, 3, flush, torch)
        ms, _ = timed_steps(p, dev, mb, steps, flush, torch)
        res[str(mb)] = {"pairs_per_s": mb * len(ms) / (sum(ms) / 1e3), "ms_per_step": float(np.mean(ms))}
        p.close()
    return res


if __name__ == "__main__":
    main()
