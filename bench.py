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

H, W, K = 480, 640, 1024
SPW = os.path.join(ROOT, "superslam_b200", "weights", "superpoint_v1.ssbw")
METRIC = "stereo frame-pairs/sec (SPx2+LG, 1024 kpts, 640x480)"
CPU_ARM = {
    "reference": "StereoFrontEnd::process over the reference's own SuperPoint / LightGlue C++ classes (compiled in place, "
                 "oracle/_ref/libref_e2e.so); their two TensorRT engines have no CPU form and are served by the fp32 torch "
                 "graphs of the oracle",
    "port": "oracle = the reference's torch graph + restated host logic (the reference itself has no CPU path, TensorRT only)",
}
WORKLOAD = ("C2: stereo pairs 640x480, K=1024, LightGlue 9 layers (seeded synthetic LightGlue weights: none ship with "
            "the reference); SuperPoint weights = reference checkpoint")

# 2*MAC counts of the dense contractions
SP_LAYER_GF = {"sp.conv1ab": 22.65 + 0.35, "sp.conv2a": 5.66, "sp.conv2b": 5.66, "sp.conv3a": 2.83, "sp.conv3b": 5.66,
               "sp.conv4a": 1.42, "sp.conv4b": 1.42, "sp.convPaDa": 5.66, "sp.convPb": 0.16, "sp.convDb": 0.63}
# LightGlue, per launch and per PAIR at N = M = 1024 (one launch covers both images of every pair)
_N = 1024
LG_LAUNCH_GF = {
    "lg.qkv": 2 * 2 * _N * 256 * 768 / 1e9, "lg.out_proj": 2 * 2 * _N * 256 * 256 / 1e9,
    "lg.ffn1": 2 * 2 * _N * 512 * 512 / 1e9, "lg.ffn2": 2 * 2 * _N * 512 * 256 / 1e9,
    "lg.qkv_cross": 2 * 2 * _N * 256 * 512 / 1e9, "lg.to_out": 2 * 2 * _N * 256 * 256 / 1e9,
    "lg.attn_self": 2 * 2 * 2 * _N * _N * 256 / 1e9,   # 2 images x (QK^T + PV)
    "lg.attn_cross": 3 * 2 * _N * _N * 256 / 1e9,      # sim once + two PV (SURVEY §8d counts sim once)
    "lg.final_proj": 2 * 2 * _N * 256 * 256 / 1e9, "lg.sim": 2 * _N * _N * 256 / 1e9,
}


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


def oracle_pair_seconds(n_iters: int, warmup: int, budget_s: float = 25.0, reference_classes: bool = False):
    """One stereo pair through the CPU oracle (fp32 torch restatement of the reference graph +
    restated host logic).  The intra-op thread count is calibrated first (8, 16, ... up to every host
    thread; the fastest wins: on a 128-thread box the oracle's many small ops run several times slower with
    128 threads than with 16), then up to `n_iters` pairs are timed, stopping early once `budget_s` seconds
    of timed work have passed.  Returns (median seconds per pair, threads used, pairs timed)."""
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
        m0, ms0 = olg.match(wlg, olg.normalize_keypoints(res[0]["xy"], W, H), res[0]["desc"],
                            olg.normalize_keypoints(res[1]["xy"], W, H), res[1]["desc"])
        q, tr, _ = ofe.dmatches(m0, ms0)
        ofe.stereo_postfilter(res[0]["xy"], res[1]["xy"], q, tr)
        return time.perf_counter() - t

    one_pair, kind = one_pair_port, "port"
    if reference_classes:   # --impl reference: the reference's own wrapper classes when they were compiled in place
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
    return float(np.median(times)), best_c, len(times), kind


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


def run_reference(args, rank: int):
    if rank != 0:
        return
    sec, threads, timed, kind = oracle_pair_seconds(max(1, args.steps), min(1, args.warmup), budget_s=90.0,
                                                    reference_classes=True)
    v = 1.0 / sec
    line = {
        "metric": METRIC, "value": v, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": {"workload": WORKLOAD, "pairs_per_step": 1,
                   "sample": "each step = one whole pair of that workload on the host cores (a bounded sample of the "
                             "64-pair device step)"},
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": threads, "kind": kind,
                         "sample": f"{timed} pair(s) timed after thread-count calibration and warm-up; " + CPU_ARM[kind]},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=64, help="stereo pairs per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    claim_stdout()

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if args.steps > 5:
            args.steps = 5  # bounded sample of the workload: whole pairs, at most 5 (and at most ~90 s)
        run_reference(args, rank)
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
    from superslam_b200.synth import synth_pair

    lib = _lib.load()
    lib.ssb_profile_enable.argtypes = [C.c_int]
    lib.ssb_profile_report.argtypes = [C.c_char_p, C.c_size_t]
    P = args.pairs
    pipe = fe.FramePairPipeline(SPW, lg_weights_path(rank), K, W, H, max_pairs=P, device=local)
    images = []
    for i in range(P):
        l, r = synth_pair(H, W, 1234 + rank * P + i)
        images += [l, r]
    dev_images = pipe.upload(images)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")  # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ----
    for _ in range(max(3, args.warmup)):
        pipe.enqueue_device(dev_images, P, H, W)
    pipe.sync()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    launches0 = fe.kernel_launch_count()
    step_ms = []
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        pipe.event_record(0)
        pipe.enqueue_device(dev_images, P, H, W)
        pipe.event_record(1)
        pipe.sync()
        step_ms.append(pipe.event_elapsed_ms(0, 1))
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

    dev_total_ms = float(sum(step_ms))
    t = torch.tensor([dev_total_ms], dtype=torch.float64, device=f"cuda:{local}")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_ms = float(t.item())
    value = world * P * args.steps / (max_ms / 1e3)

    # ---- end to end through the public call, host buffers in / out ----
    # Streaming API (ssb_fe_submit / ssb_fe_collect): every step uploads its 2P images from pinned host memory
    # and reads its results back to the host; the upload of step i+1 is in flight while step i computes.
    pinned = [torch.from_numpy(im).pin_memory() for im in images]
    images = [t.numpy() for t in pinned]
    for _ in range(2):
        pipe.process(images)
    pipe.submit(images)
    for _ in range(6):           # both image buffers seen three times: eager, capture, replay
        pipe.submit(images)
        pipe.collect()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pipe.submit(images)
        res = pipe.collect()     # results of the previous step; one step stays in flight across the loop
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    res = pipe.collect()
    # the same through the synchronous call (upload -> kernels -> read-back, nothing overlapped)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pipe.process(images)
    e2e_sync_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s, e2e_sync_s], dtype=torch.float64, device=f"cuda:{local}")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * P * args.steps / float(t[0].item())
    e2e_sync_value = world * P * args.steps / float(t[1].item())
    h2d = 2 * P * H * W
    d2h = sum(v.nbytes for v in res.values())

    # result gather (the only collective on the path): per-rank match counts
    matches = torch.tensor([int((out["matches0"] >= 0).sum()), int(out["has_depth"].sum()), int(out["count"].sum())],
                           dtype=torch.int64, device=f"cuda:{local}")
    gathered = [matches]
    if dist is not None:
        gathered = [torch.zeros_like(matches) for _ in range(world)]
        dist.all_gather(gathered, matches)

    # ... and the result gather proper: this rank's pairs of the last step as fixed-size padded records
    # (pair s of rank r is pair r + world * s of the round-robin stream), one all_gather over NCCL, unpacked on every rank
    try:
        from superslam_b200 import sharding

        rec = sharding.pack_records([rank + world * s for s in range(P)], out["count"].tolist(), out["matches0"],
                                    out["mscores0"], out["has_depth"], K, P)
        allrec = sharding.gather_records(rec, dist, device=f"cuda:{local}") if dist is not None else rec[None]
        got = sharding.unpack_records(allrec, K)
        records = {"pairs_gathered": len(got), "record_bytes_per_rank": int(rec.nbytes),
                   "matches_in_records": int(sum(int((r["matches0"][:r["n_left"]] >= 0).sum()) for r in got.values()))}
    except Exception as e:  # never at the cost of the headline line
        records = {"error": str(e)}

    if rank == 0:
        # roofline of the dominant tensor-core kernel
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained"
        total_prof_ms = sum(ms for _, ms in prof.values()) or 1.0
        dom = max(prof.items(), key=lambda kv: kv[1][1])[0] if prof else None
        roof = None
        if dom is not None:
            cnt, ms = prof[dom]
            avg_ms = ms / max(1, cnt)
            if dom in SP_LAYER_GF or dom in LG_LAUNCH_GF:
                # algorithmic GFLOP per launch: SuperPoint layers see 2*P images, LightGlue launches P pairs
                gf = SP_LAYER_GF[dom] * 2 * P if dom in SP_LAYER_GF else LG_LAUNCH_GF[dom] * P
                traffic = None
                try:  # DRAM bytes per launch from the committed ncu capture (per image / per pair), times this run's batch
                    tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_r01.json"))).get(dom)
                    if tr:
                        traffic = int(tr["bytes"] * (2 * P if tr["per"] == "image" else P))
                except Exception:
                    pass
                roof = {"kernel": dom, "bound": "tensor", "achieved": gf / avg_ms, "peak": peak_tf, "unit": "TFLOP/s",
                        "frac": gf / avg_ms / peak_tf, "traffic": traffic,
                        "traffic_source": "profiles/ncu_traffic_r01.json (ncu dram__bytes_read.sum + dram__bytes_write.sum per image/pair at 64 pairs/step, x this run's batch)",
                        "avg_launch_ms": avg_ms,
                        "share_of_step": ms / total_prof_ms, "peak_source": peak_src,
                        "algorithmic_gflop_per_launch": gf}
            else:
                roof = {"kernel": dom, "bound": "hbm", "achieved": None, "peak": float(peaks.get("hbm_gbs", 6650.0)),
                        "unit": "GB/s", "frac": None, "traffic": None, "avg_launch_ms": avg_ms,
                        "share_of_step": ms / total_prof_ms, "peak_source": peak_src}
        shares = {k: round(v[1] / total_prof_ms, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:12]}
        tensor_kernels = {}
        for name, (cnt, ms) in prof.items():
            gfl = SP_LAYER_GF[name] * 2 * P if name in SP_LAYER_GF else LG_LAUNCH_GF.get(name, 0) * P
            if gfl > 0 and ms > 0:
                tensor_kernels[name] = {"avg_launch_ms": round(ms / cnt, 5), "tflops": round(gfl / (ms / cnt), 1),
                                        "frac_of_peak": round(gfl / (ms / cnt) / peak_tf, 4)}
        # the HBM-bound pieces: algorithmic bytes per launch / measured launch time against the measured copy
        # bandwidth.  Bytes per unit as stated in DESIGN.md section 3: NMS reads the fp32 heat map once (4 H' W' per image),
        # the gather moves one 512-byte grid row in and one descriptor row out per keypoint (K * 1024 per image), the
        # assignment sweeps read sim and sim^T once each (2 * K * K * 4 per pair).
        hbm_kernels = {}
        try:
            peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
            hc8, wc8 = (H // 8) * 8, (W // 8) * 8
            per_launch_bytes = {"sp.nms": 4 * hc8 * wc8 * 2 * P, "sp.gather": K * 1024 * 2 * P,
                                "lg.lse": 2 * K * K * 4 * P, "lg.argmax": 2 * K * K * 4 * P}
            for name, nbytes in per_launch_bytes.items():
                if name in prof and prof[name][1] > 0:
                    cnt, ms = prof[name]
                    gbs = nbytes / (ms / cnt) / 1e6
                    hbm_kernels[name] = {"avg_launch_ms": round(ms / cnt, 5), "algorithmic_mb_per_launch": round(nbytes / 1e6, 2),
                                         "gb_per_s": round(gbs, 1), "frac_of_peak": round(gbs / peak_gbs, 4)}
        except Exception as e:  # a reporting extra must never cost the headline line
            hbm_kernels = {"error": str(e)}
        eigen = None
        if world == 1:
            try:
                eigen = bench_eigenplaces(lib, local)
            except Exception as e:  # the headline line must not depend on the "next" row
                eigen = {"error": str(e)}
        latency = None
        if world == 1:
            try:
                latency = bench_latency(local, rank)
            except Exception as e:
                latency = {"error": str(e)}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            sec, threads, timed, kind = oracle_pair_seconds(5, 1, budget_s=20.0)
            cpu = {"value": 1.0 / sec, "unit": "pairs/s", "cores": threads, "kind": kind,
                   "sample": f"{timed} pair(s) of the same C2 workload after warm-up and thread-count calibration; " + CPU_ARM[kind]}
        line = {
            "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": max_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16 (fp32 accumulate)", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "pairs_per_step_per_gpu": P, "l2": "flushed between steps (256 MiB memset)",
                       "timing": "CUDA events per step on the pipeline stream (CUDA-graph replay), max over ranks; "
                                 "roofline/kernel shares from an eager re-run of the same steps with an event "
                                 "after every kernel"},
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "FramePairPipeline.submit/collect (streaming: H2D of step i+1 under the kernels of step i)",
                    "synchronous_process_value": e2e_sync_value},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "kernel_time_shares": shares,
            "tensor_kernels": tensor_kernels,
            "hbm_kernels": hbm_kernels,
            "kernel_ms_per_step": {k: round(v[1] / args.steps, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])},
            "cpu_baseline": cpu,
            "latency_single_pair": latency,
            "eigenplaces": eigen,
            "wall_s_timed_region": wall,
            "results": {"per_rank_[matches,has_depth,keypoints]": [g.tolist() for g in gathered],
                        "gathered_records": records},
        }
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
