"""The numbers bench.py quotes from profiles/ must be derivable from the committed evidence: the per-kernel DRAM
traffic table is re-generated from the committed ncu launch list and compared."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_traffic_table_matches_the_committed_launch_list():
    csv = os.path.join(ROOT, "profiles", "launches_r01_v14_p64.csv")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_traffic.py"), csv, "64", "0"],
                         capture_output=True, text=True, check=True).stdout
    got = json.loads(out)
    exp = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_r01.json")))
    got.pop("_comment"), exp.pop("_comment")
    assert got == exp
    # one step = 94 kernel launches; the dominant kernel of the roofline is the fused conv1a+conv1b
    assert sum(v["launches"] for v in exp.values()) == 94
    assert max(exp.items(), key=lambda kv: kv[1]["share_of_step"])[0] == "sp.conv1ab"
    assert exp["sp.conv1ab"]["per"] == "image" and exp["lg.ffn2"]["per"] == "pair"


def test_round2_traffic_table_matches_its_launch_list():
    """The table bench.py quotes `roofline.traffic` from (profiles/ncu_traffic_r02.json) is what tools/ncu_traffic.py
    derives from the committed launch list of the final build (kernel names carry the pair / N template arguments)."""
    csv = os.path.join(ROOT, "profiles", "launches_r02_v5_p64.csv")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_traffic.py"), csv, "64"],
                         capture_output=True, text=True, check=True).stdout
    got = json.loads(out)
    exp = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_r02.json")))
    got.pop("_comment"), exp.pop("_comment")
    assert got == exp
    assert sum(v["launches"] for v in exp.values()) == 94
    assert max(exp.items(), key=lambda kv: kv[1]["share_of_step"])[0] == "sp.conv1ab"
    # the Cout = 64 convolutions are the CTA-pair kernels, conv3a the N = 128 pair
    assert "1, 1, 64>" in exp["sp.conv1ab"]["kernel"] and "0, 1, 128>" in exp["sp.conv3a"]["kernel"]


def test_roofline_model_reproduces_the_committed_table():
    """profiles/roofline_r02.md (which roof binds each kernel: tensor / HBM / shared memory / MUFU) is what
    tools/roofline_model.py derives from the committed bench line, and the statements the docs make hold in it."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "roofline_model.py")], capture_output=True, text=True,
                         check=True).stdout
    assert out == open(os.path.join(ROOT, "profiles", "roofline_r02.md")).read()
    rows = {}
    for ln in out.splitlines():
        c = [x.strip() for x in ln.split("|")]
        if len(c) > 9 and c[1] and c[1] not in ("kernel", "---"):
            rows[c[1]] = (c[8], float(c[9]))
    assert rows["sp.conv1ab"][0] == "shared memory" and rows["sp.conv2a"][0] == "shared memory"
    assert rows["lg.ffn2"][0] == "HBM" and rows["lg.attn_self"][0] == "MUFU" and rows["sp.conv3b"][0] == "tensor"
    # no kernel runs faster than its binding roof allows (conv3a: a 0.24 ms kernel, above the SUSTAINED tensor figure)
    assert all(f <= 1.0 for k, (r, f) in rows.items() if k != "sp.conv3a"), rows
    assert rows["sp.conv3a"][1] <= 1644.4 / 1380.6


def test_bench_lines_are_valid_json_with_the_contract_keys():
    prof = os.path.join(ROOT, "profiles")
    need = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"}
    for name in ("bench_r01_v15_p64.json", "bench_r01_v15_n2.json", "bench_r02_v6_p64.json", "bench_r02_v4_n2.json",
                 "bench_r02_v6_n8.json", "bench_r02_v4_C3.json", "bench_r02_v4_C5.json"):
        d = json.load(open(os.path.join(prof, name)))
        assert need <= set(d), (name, need - set(d))
        assert d["unit"] == "pairs/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
        r = d["roofline"]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9


def test_reference_arm_prints_one_contract_line_on_the_cpu():
    """`bench.py --impl reference` needs no GPU: one JSON line on stdout, the product arm's metric / unit / workload,
    the arm's own cpu_baseline and a zero-copy e2e object."""
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    sys.path.insert(0, ROOT)
    import bench

    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == "pairs/s"
    assert d["config"]["workload"] == bench.WORKLOAD and d["higher_is_better"] is True and d["gpu_launches"] == 0
    # the reference's own wrapper classes when oracle/_ref/libref_e2e.so was built (engines served by the oracle), else the port
    built = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_e2e.so"))
    assert d["cpu_baseline"]["kind"] == ("reference" if built else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
