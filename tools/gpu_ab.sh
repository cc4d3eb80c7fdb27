# usage: tools/gpu_ab.sh [tag]   (run under gpurun): GPU tests, then bench A/B over env toggles
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
show() { python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
k=d['kernel_ms_per_step']
print(sys.argv[1], round(d['value'],1), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'])
print('  ', {n:v for n,v in list(k.items())[:14]})
PY
}
timeout 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_new.json 2>gpurun_out/bench_new.err; show gpurun_out/bench_new.json
SSB_FA_PTMEM=0 timeout 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_nopt.json 2>>gpurun_out/bench_new.err; show gpurun_out/bench_nopt.json
SSB_LG_FOLD_OUT=0 timeout 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_nofold.json 2>>gpurun_out/bench_new.err; show gpurun_out/bench_nofold.json
tail -3 gpurun_out/bench_new.err
