#!/usr/bin/env python3
"""Condense an `ncu --set full` report into the few per-launch numbers the design notes quote.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/ncu_<name>_rNN.txt

Needs the `ncu` CLI (reads the report with `--page raw --csv`); no GPU required.
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__cycles_elapsed.avg", "SM cycles"),
    ("sm__inst_executed.sum", "warp instructions"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue slots busy"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe (hmma) active"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    col = {c: i for i, c in enumerate(head)}
    stall_cols = [(c, i) for c, i in col.items()
                  if c.startswith("smsp__average_warps_issue_stalled_") and c.endswith("_per_issue_active.ratio")]
    print(f"# {rep}: {len(rows) - 2} launch(es), ncu --set full --clock-control none")
    for n, row in enumerate(rows[2:]):
        print(f"\n[{n}] {row[col['Kernel Name']][:110]}")
        for key, label in METRICS:
            if key in col and row[col[key]] != "":
                print(f"    {label:28s} {row[col[key]]:>16s} {units[col[key]]}")
        stalls = []
        for c, i in stall_cols:
            try:
                stalls.append((float(row[i]), c[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
        stalls.sort(reverse=True)
        print("    top stalls (warps per issue) " + ", ".join(f"{name} {v:.2f}" for v, name in stalls[:5]))


if __name__ == "__main__":
    main()
