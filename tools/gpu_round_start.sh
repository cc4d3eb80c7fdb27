# usage (under gpurun, about 10 GPU-minutes):  gpurun --timeout 1500 -- 'bash tools/gpu_round_start.sh r02'
# The first device call of a round: everything whose result shapes the plan, in one box.
#   1. all GPU tests WITHOUT -x (every failure is listed, incl. the tests written after the last device run)
#   2. smoke()
#   3. stage-by-stage parity report at the C2 size incl. the matches0 disagreement margins   -> diag_c2_<tag>.txt
#   4. the default bench                                                                        -> bench_<tag>.json
#   5. ncu launch list with DRAM bytes of one bench step                                        -> launches_<tag>.csv
tag=${1:-r02}
mkdir -p gpurun_out
echo "== 1. pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/pytest_$tag.log
echo "== 2. smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== 3. diag"; timeout 600 python tests/gpu_diag.py --size 480x640 --k 1024 > gpurun_out/diag_c2_$tag.txt 2>&1; tail -40 gpurun_out/diag_c2_$tag.txt
echo "== 4. bench"; timeout 600 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python - "$tag" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/bench_{sys.argv[1]}.json"))
print(round(d["value"], 1), round(d["e2e"]["value"], 1), d["clocks"], d["roofline"]["kernel"], round(d["roofline"]["frac"], 3))
print({k: v for k, v in list(d["kernel_ms_per_step"].items())[:16]})
print(d.get("hbm_kernels")); print(d["results"].get("gathered_records")); print(d.get("cpu_baseline"))
PY
echo "== 5. ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/launches_$tag.csv python bench.py --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/ncu_$tag.log 2>&1
python tools/ncu_traffic.py gpurun_out/launches_$tag.csv 64 > gpurun_out/ncu_traffic_$tag.json 2>> gpurun_out/ncu_$tag.log && echo "traffic table written"
