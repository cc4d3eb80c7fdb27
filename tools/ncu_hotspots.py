#!/usr/bin/env python3
"""Per-instruction warp-stall hot spots of one launch in an `ncu --set full --import-source on` report.

    python tools/ncu_hotspots.py gpurun_out/full.ncu-rep <launch index> [top N]

Reads the report with `ncu --page source --csv` (no GPU needed) and prints the N SASS instructions with the
most stall samples, in program order, with their two dominant stall reasons.
"""
import csv
import io
import subprocess
import sys


def main():
    rep, launch = sys.argv[1], int(sys.argv[2])
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(launch),
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    print("#", rows[0][1] if len(rows[0]) > 1 else rows[0])
    hdr = rows[1]
    col = {c: i for i, c in enumerate(hdr)}
    data = []
    for r in rows[2:]:          # the SASS view comes first; stop at the next section header
        if len(r) != len(hdr) or r[col["# Samples"]] == "# Samples":
            break
        data.append(r)
    tot = sum(int(r[col["# Samples"]]) for r in data) or 1
    stalls = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    print(f"# {tot} samples over {len(data)} instructions")
    by_reason = {c: sum(int(r[col[c]]) for r in data) for c in stalls}
    print("# by reason:", ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(by_reason.items(), key=lambda kv: -kv[1])[:8]))
    idx = sorted(range(len(data)), key=lambda i: -int(data[i][col["# Samples"]]))[:top]
    for i in sorted(idx):
        r = data[i]
        s = int(r[col["# Samples"]])
        st = sorted(((int(r[col[c]]), c[6:]) for c in stalls), reverse=True)[:2]
        print(f"{i:5d} {r[col['Source']].strip()[:72]:72s} {100 * s / tot:5.1f}%  " + ", ".join(f"{n} {v}" for v, n in st if v))


if __name__ == "__main__":
    main()
