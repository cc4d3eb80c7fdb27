#!/usr/bin/env python3
"""Which roof binds each kernel of the C2 step: tensor, HBM, shared memory or the MUFU?

    python tools/roofline_model.py [profiles/bench_r02_v6_p64.json] > profiles/roofline_r02.md

For every kernel of one 64-pair step the script states the ALGORITHMIC work per launch in four currencies - tensor-core
FLOPs (2 * MAC), HBM bytes (every tensor read once and written once), shared-memory bytes (what the kernel's structure
moves through the 128 B/clk port of an SM: TMA writes, tcgen05.mma operand reads, staging stores and their TMA-store
reads) and exponentials - turns each into a time at the measured peaks (the sustained bf16 TFLOP/s and copy GB/s the bench line was
scored against; the shared-memory port and the MUFU at the SM clock the bench line reports), and puts the measured launch
time (bench.py, CUDA events) next to the largest of them.  No GPU needed: it only reads committed files.

The per-tile shared-memory byte counts are the ones derived in profiles/README.md (round 2) and csrc/conv_pipe.cuh /
umma_core.cuh; they are written out below so that they can be checked against the code.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMS = 148
SMEM_B_PER_CLK = 128          # per SM
MUFU_PER_CLK = 16             # ex2 per clk per SM (tools/mufu_probe.cu: 15.9 measured)

H, W, K, P = 480, 640, 1024, 64
IMG = 2 * P
ROWS = IMG * K                # LightGlue rows per launch


def conv_tiles(h, w):
    return IMG * ((h + 15) // 16) * ((w + 15) // 16)


def kernels():
    """name -> dict(gf, hbm_mb, smem_mb, exps, launches) per LAUNCH (launches = per step)."""
    k = {}
    px = H * W
    # ---- SuperPoint (conv_pipe.cuh, CTA pairs: per tile 72 MMAs read 4 KB of A + 1 KB of B each) ----
    t1 = conv_tiles(H, W)
    k["sp.conv1ab"] = dict(gf=IMG * 2 * px * 9 * (1 * 64 + 64 * 64) / 1e9, hbm_mb=IMG * (px + px // 4 * 64 * 2) / 1e6,
                           # conv1b operands 72 x 5 KB, conv1a operands 12 x 5 KB, im2col 10.4 + 6.6 KB, halo stores 41.5 KB,
                           # pooled output staged and read back 2 x 8 KB
                           smem_mb=t1 * (72 * 5120 + 12 * 5120 + 17000 + 41472 + 16384) / 1e6, launches=1)
    t2 = conv_tiles(H // 2, W // 2)
    for name, pooled in (("sp.conv2a", False), ("sp.conv2b", True)):
        out_b = (H // 2) * (W // 2) * 64 * 2 // (4 if pooled else 1)
        k[name] = dict(gf=IMG * 2 * (px // 4) * 9 * 64 * 64 / 1e9, hbm_mb=IMG * ((px // 4) * 64 * 2 * 1.27 + out_b) / 1e6,
                       # operands 72 x 5 KB, halo box written by TMA 41.5 KB, output staged + read back
                       smem_mb=t2 * (72 * 5120 + 41472 + 2 * (8192 if pooled else 32768)) / 1e6, launches=1)
    t3 = conv_tiles(H // 4, W // 4)
    k["sp.conv3a"] = dict(gf=IMG * 2 * (px // 16) * 9 * 64 * 128 / 1e9, hbm_mb=IMG * (px // 16) * (64 * 2 * 1.27 + 128 * 2) / 1e6,
                          # N = 128 pair: 72 MMAs x (4 KB of A + 2 KB of B), halo 41.5 KB, 64 KB of output staged + read back
                          smem_mb=t3 * (72 * 6144 + 41472 + 2 * 65536) / 1e6, launches=1)
    # Cin = 128 layers (conv_stream.cuh, one CTA per SM): per 16 x 16 tile and 128-channel slice 144 N = 128 MMAs read
    # 4 KB of A + 4 KB of B each (64 cycles of math, 64 cycles of port time), the weights stream through shared memory
    # once per tile (18 stages of 16 KB), two 41.5 KB halo slabs, the output staged and read back
    def stream(h, w, slices, pooled):
        tiles = conv_tiles(h, w) * slices
        out_b = 256 * 128 * 2 // (4 if pooled else 1)
        return tiles * (144 * 8192 + 18 * 16384 + 2 * 41472 + 2 * out_b) / 1e6
    k["sp.conv3b"] = dict(gf=IMG * 2 * (px // 16) * 9 * 128 * 128 / 1e9, hbm_mb=IMG * (px // 16) * 128 * 2 * (1.27 + 0.25) / 1e6,
                          smem_mb=stream(H // 4, W // 4, 1, True), launches=1)
    for name in ("sp.conv4a", "sp.conv4b"):
        k[name] = dict(gf=IMG * 2 * (px // 64) * 9 * 128 * 128 / 1e9, hbm_mb=IMG * (px // 64) * 128 * 2 * 2.27 / 1e6,
                       smem_mb=stream(H // 8, W // 8, 1, False), launches=1)
    k["sp.convPaDa"] = dict(gf=IMG * 2 * (px // 64) * 9 * 128 * 512 / 1e9, hbm_mb=IMG * (px // 64) * (128 * 2 * 1.27 + 512 * 2) / 1e6,
                            smem_mb=stream(H // 8, W // 8, 4, False), launches=1)
    k["sp.convPb"] = dict(gf=IMG * 2 * (px // 64) * 256 * 65 / 1e9, hbm_mb=IMG * ((px // 64) * 256 * 2 + px * 4) / 1e6, launches=1)
    k["sp.convDb"] = dict(gf=IMG * 2 * (px // 64) * 256 * 256 / 1e9, hbm_mb=IMG * (px // 64) * 256 * 2 * 2 / 1e6, launches=1)
    k["sp.nms"] = dict(hbm_mb=IMG * px * 4 / 1e6, launches=1)
    k["sp.gather"] = dict(hbm_mb=IMG * K * 1024 / 1e6, launches=1)
    # ---- LightGlue (umma_core.cuh; pair mode: per K chunk a CTA writes and reads 16 + 16 KB) ----
    def lin(kdim, n, out_mb, extra_in_mb=0.0, pair=True):
        tiles = ROWS // 128 * (n // 256)
        chunk = (16384 + (16384 if pair else 32768)) * 2          # TMA write + MMA read per 64-wide K chunk
        return dict(gf=2 * ROWS * kdim * n / 1e9, hbm_mb=ROWS * kdim * 2 / 1e6 + extra_in_mb + out_mb,
                    smem_mb=tiles * (kdim // 64 * chunk + 2 * 128 * 256 * 2) / 1e6)
    k["lg.qkv"] = dict(lin(256, 768, ROWS * 768 * 2 / 1e6), launches=9)
    k["lg.qkv_cross"] = dict(lin(256, 512, ROWS * 512 * 2 / 1e6), launches=9)
    k["lg.ffn1"] = dict(lin(512, 512, ROWS * 512 * 2 / 1e6, pair=False), launches=18)
    k["lg.ffn2"] = dict(lin(512, 256, ROWS * 256 * (4 + 2) / 1e6, extra_in_mb=ROWS * 256 * 4 / 1e6), launches=18)
    k["lg.final_proj"] = dict(lin(256, 256, ROWS * 768 * 2 * 2 / 1e6), launches=1)
    blocks = IMG * 4 * (K // 128) * (K // 128)                     # 128 x 128 logit blocks per attention launch
    for name, flops_per_block in (("lg.attn_self", 4 * 128 * 128 * 64), ("lg.attn_cross", 3 * 128 * 128 * 64)):
        k[name] = dict(gf=blocks * flops_per_block / 1e9, exec_gf=blocks * 4 * 128 * 128 * 64 / 1e9,
                       hbm_mb=ROWS * 256 * 2 * 4 / 1e6, exps=blocks * 128 * 128,
                       # per block: S operands 4 x 8 KB, V operands 8 x 2 KB, half a K + V block of TMA writes (two tiles share them)
                       smem_mb=blocks * (4 * 8192 + 8 * 2048 + 16384) / 1e6, launches=9)
    k["lg.sim"] = dict(gf=P * 2 * K * K * 768 / 1e9, hbm_mb=P * (2 * K * 768 * 2 + K * K * 4) / 1e6, launches=1)
    k["lg.lse"] = dict(hbm_mb=P * K * K * 4 / 1e6, launches=1)
    k["lg.argmax"] = dict(hbm_mb=P * K * K * 4 / 1e6, launches=1)
    return k


def main():
    bench = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "bench_r02_v6_p64.json")
    d = json.load(open(bench))
    # the peaks the bench line itself was scored against (MEASURED_PEAKS.json of the box that produced it)
    tf = float(d["roofline"]["peak"]) if d["roofline"].get("unit") == "TFLOP/s" else 1380.6
    hk = next(iter(d.get("hbm_kernels", {}).values()), None)
    gbs = round(hk["gb_per_s"] / hk["frac_of_peak"]) if hk else 6538.0
    mhz = d["clocks"]["sm_mhz"] or 1750.0
    smem_gbs = SMS * SMEM_B_PER_CLK * mhz * 1e6 / 1e9
    mufu_per_s = SMS * MUFU_PER_CLK * mhz * 1e6
    ms = d["kernel_ms_per_step"]
    print(f"# Which roof binds each kernel (C2, 64 pairs per step; {os.path.basename(bench)})\n")
    print(f"Peaks: tensor {tf:.0f} TFLOP/s (measured sustained bf16), HBM {gbs:.0f} GB/s (measured copy), shared memory "
          f"{smem_gbs / 1e3:.1f} TB/s (148 SMs x 128 B/clk at the {mhz:.0f} MHz of this run), MUFU "
          f"{mufu_per_s / 1e12:.2f} T exp/s.  Times are per launch in ms; `frac` = largest bound / measured.  Generated by "
          f"`tools/roofline_model.py`.\n")
    print("| kernel | launches | measured | tensor | HBM | shared memory | MUFU | binding roof | frac |")
    print("|---|---|---|---|---|---|---|---|---|")
    total_meas = total_bound = 0.0
    rows = {}
    for name, w in kernels().items():
        if name not in ms:
            continue
        n = w["launches"]
        meas = ms[name] / n
        b = {"tensor": w.get("exec_gf", w.get("gf", 0.0)) / tf, "HBM": w.get("hbm_mb", 0.0) / gbs,
             "shared memory": w.get("smem_mb", 0.0) / smem_gbs, "MUFU": w.get("exps", 0.0) / mufu_per_s * 1e3}
        roof = max(b, key=b.get)
        frac = b[roof] / meas
        rows[name] = (roof, frac)
        total_meas += ms[name]
        total_bound += b[roof] * n
        f = lambda v: f"{v:.3f}" if v > 0 else "-"
        print(f"| {name} | {n} | {meas:.3f} | {f(b['tensor'])} | {f(b['HBM'])} | {f(b['shared memory'])} | {f(b['MUFU'])} | {roof} | {frac:.2f} |")
    print(f"\nSum over these kernels: measured {total_meas:.2f} ms per step, sum of the binding bounds {total_bound:.2f} ms "
          f"({total_bound / total_meas:.2f}).  Cross attention is counted with the FLOPs it executes (QK^T per direction); "
          f"the algorithmic count of `bench.py` takes QK^T once per pair.")
    return rows


if __name__ == "__main__":
    main()
