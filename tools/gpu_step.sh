# usage (under gpurun):  bash tools/gpu_step.sh <tag> [what...]     what = tests diag bench ab configs ncu  (default: tests diag bench)
# One development step on the GPU box: every part is bounded by its own timeout and none blocks the next.
tag=${1:-step}; shift
what=${@:-tests diag bench}
mkdir -p gpurun_out
for w in $what; do
case $w in
tests) echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/pytest_$tag.log ;;
tests_all) echo "== pytest -m gpu (no -x)"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee gpurun_out/pytest_$tag.log ;;
lgtests) echo "== pytest lightglue + pipeline"; timeout 600 python -m pytest tests/test_gpu_lightglue.py tests/test_gpu_pipeline.py -q -x 2>&1 | tail -25 | tee gpurun_out/pytest_$tag.log ;;
smoke) echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ;;
diag) echo "== diag"; timeout 600 python tests/gpu_diag.py --size 480x640 --k 1024 > gpurun_out/diag_c2_$tag.txt 2>&1; grep -E "lg x32 final|lg matches0|disagreements|e2e\[|margin" gpurun_out/diag_c2_$tag.txt ;;
bench) echo "== bench"; timeout 600 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -3 gpurun_out/bench_$tag.err
python - "$tag" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/bench_{sys.argv[1]}.json"))
print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["clocks"], d["roofline"]["kernel"], round(d["roofline"]["frac"], 3), "pipeline frac", round(d["pipeline_roofline"]["frac"], 3))
print(d["kernel_ms_per_step"]); print("spread", d["value_spread_per_step"], "e2e repeats", d["e2e"]["repeats"])
print("live", d.get("live_pipeline")); print("latency", d.get("latency_single_pair")); print("records", d["results"]["gathered_records"])
PY
;;
ab) # ABS="ENV=1 ENV2=x ..." (default: SSB_LG_PAIR=0): one short bench per entry with that environment assignment
for ab in ${ABS:-SSB_LG_PAIR=0}; do
  echo "== A/B: $ab"; env $ab timeout 300 python bench.py --no-cpu-baseline --no-extras --steps 10 > gpurun_out/bench_${tag}_ab.json 2> gpurun_out/bench_${tag}_ab.err
  python - "$tag" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/bench_{sys.argv[1]}_ab.json"))
print("   value", round(d["value"], 1), d["clocks"]["sm_mhz"], "MHz;", {k: v for k, v in list(d["kernel_ms_per_step"].items())[:6]})
PY
done ;;
configs) for c in C1 C3 C4 C5; do echo "== bench --config $c"; timeout 600 python bench.py --config $c --steps 5 --no-cpu-baseline > gpurun_out/bench_${tag}_$c.json 2> gpurun_out/bench_${tag}_$c.err; tail -2 gpurun_out/bench_${tag}_$c.err
python - "$tag" "$c" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/bench_{sys.argv[1]}_{sys.argv[2]}.json"))
    print(sys.argv[2], "value", round(d["value"], 1), d["unit"], "e2e", round(d["e2e"]["value"], 1), "pipeline frac", round(d["pipeline_roofline"]["frac"], 3), "attn share of LG flops", d["pipeline_roofline"]["attention_frac_of_lightglue_flops"], d["config"]["keypoints_per_image"], "sweep", d.get("micro_batch_sweep"))
    print({k: v for k, v in list(d["kernel_ms_per_step"].items())[:10]})
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
done ;;
trace) # TRACES="lg.ffn1 lg.qkv ...": per-tile time stamps of CTA 0 (MMA warp / epilogue warp) of the first launch with that label
for lb in ${TRACES:-lg.ffn1 lg.qkv lg.ffn2}; do
  echo "== core trace $lb"; SSB_CORE_TRACE=$lb timeout 300 python bench.py --no-cpu-baseline --no-extras --steps 1 --warmup 1 2>&1 >/dev/null | grep -A14 "core trace" | head -16
done ;;
ncufull) # NCU_KERNELS="name:regex ..." one full capture (source counters included) of the 3rd launch matching each regex
for spec in ${NCU_KERNELS:-attn:flash_attention qkv:EpiQkvRope ffn1:EpiLnGelu}; do
  name=${spec%%:*}; re=${spec#*:}
  echo "== ncu --set full: $name ($re)"
  timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:$re --launch-skip 2 --launch-count 1 -f \
      -o gpurun_out/full_${tag}_$name python bench.py --no-cpu-baseline --no-extras --steps 1 --warmup 1 > gpurun_out/ncufull_${tag}_$name.log 2>&1
  ls -la gpurun_out/full_${tag}_$name.ncu-rep 2>/dev/null | awk '{print $5, $9}'
done ;;
ncu) echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/launches_$tag.csv python bench.py --no-cpu-baseline --no-extras --steps 1 --warmup 1 > gpurun_out/ncu_$tag.log 2>&1
python tools/ncu_traffic.py gpurun_out/launches_$tag.csv 64 > gpurun_out/ncu_traffic_$tag.json 2>> gpurun_out/ncu_$tag.log && echo "traffic table written" ;;
esac
done
