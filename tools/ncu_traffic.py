#!/usr/bin/env python3
"""Per-kernel launch time and DRAM traffic of ONE bench step from an ncu launch list.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --csv --log-file gpurun_out/launches.csv python bench.py --no-cpu-baseline --steps 1 --warmup 1
    python tools/ncu_traffic.py gpurun_out/launches.csv <pairs per step> [block index] > profiles/ncu_traffic_rNN.json

The pair pipeline launches a fixed sequence of kernels per step (SuperPoint: 10 convolution launches, memset-free
NMS / select / gather; LightGlue: prepare, 9 x 8 block launches, 7 assignment launches; post-filter).  The
script cuts the launch list into steps at every conv1a+conv1b launch, labels the launches of the chosen step
by position and reports, per label, launches, mean duration and DRAM bytes per image (SuperPoint) or per pair
(LightGlue) - the unit bench.py multiplies by its own batch for roofline.traffic.
"""
import csv
import re
import json
import sys
from collections import OrderedDict

SP = ["sp.conv1ab", "sp.conv2a", "sp.conv2b", "sp.conv3a", "sp.conv3b", "sp.conv4a", "sp.conv4b", "sp.convPaDa",
      "sp.convPb", "sp.convDb", "sp.nms", "sp.select", "sp.gather"]
LAYER = ["lg.qkv", "lg.attn_self", "lg.ffn1", "lg.ffn2", "lg.qkv_cross", "lg.attn_cross", "lg.ffn1", "lg.ffn2"]
TAIL = ["lg.final_proj", "lg.matchability", "lg.sim", "lg.simT", "lg.lse", "lg.argmax", "lg.mutual", "fe.postfilter"]
LABELS = SP + ["lg.prepare"] + LAYER * 9 + TAIL


def main():
    path, pairs = sys.argv[1], int(sys.argv[2])
    block = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.reader(lines)
    head = next(rd)
    col = {c: i for i, c in enumerate(head)}
    launches = OrderedDict()
    for r in rd:
        k = int(r[col["ID"]])
        e = launches.setdefault(k, {"name": r[col["Kernel Name"]]})
        e[r[col["Metric Name"]]] = float(r[col["Metric Value"]].replace(",", ""))
    seq = list(launches.values())
    starts = [i for i, e in enumerate(seq) if re.search(r"conv_pipe_kernel<ssb::EpiConvRelu, (\(bool\))?1", e["name"])]   # the fused first layer
    if block >= len(starts):
        raise SystemExit(f"only {len(starts)} steps in the list")
    a = starts[block]
    b = starts[block + 1] if block + 1 < len(starts) else len(seq)
    step = [e for e in seq[a:b] if "memset" not in e["name"].lower()]
    if len(step) < len(LABELS):
        raise SystemExit(f"step has {len(step)} launches, expected {len(LABELS)}")
    out = OrderedDict()
    out["_comment"] = (f"one step of bench.py at {pairs} pairs/step ({2 * pairs} images), launches {a}..{b - 1} of {path}: "
                       "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                       "--clock-control none (serialised, cold-ish caches: compare shares, not absolutes).  bytes = DRAM "
                       "read + write per launch divided by the images (sp.*) or pairs (lg.*, fe.*) of the step.")
    total_ns = sum(e.get("gpu__time_duration.sum", 0.0) for e in step[:len(LABELS)])
    for lab, e in zip(LABELS, step):
        d = out.setdefault(lab, {"per": "image" if lab.startswith("sp.") else "pair", "launches": 0, "ns": 0.0,
                                 "read": 0.0, "write": 0.0, "kernel": e["name"][:60]})
        d["launches"] += 1
        d["ns"] += e.get("gpu__time_duration.sum", 0.0)
        d["read"] += e.get("dram__bytes_read.sum", 0.0)
        d["write"] += e.get("dram__bytes_write.sum", 0.0)
    for lab, d in out.items():
        if lab == "_comment":
            continue
        units = (2 * pairs if d["per"] == "image" else pairs) * d["launches"]
        d["read"] = int(d["read"] / units)
        d["write"] = int(d["write"] / units)
        d["bytes"] = d["read"] + d["write"]
        d["share_of_step"] = round(d["ns"] / total_ns, 4)
        d["avg_launch_us"] = round(d["ns"] / d["launches"] / 1e3, 2)
        del d["ns"]
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
