# usage (under gpurun): bash tools/gpu_final.sh <tag>   GPU tests, smoke, default bench -> gpurun_out/bench_<tag>.json
tag=${1:-final}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python - "$tag" <<'PY'
import json,sys
d=json.load(open(f'gpurun_out/bench_{sys.argv[1]}.json'))
k=d['kernel_ms_per_step']
print(round(d['value'],1), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'], d['roofline']['frac'])
print({n:v for n,v in list(k.items())[:14]})
print(d.get('latency_single_pair'))
print(d.get('cpu_baseline'))
PY
