# usage (under gpurun): bash tools/gpu_ab.sh [ENV=VAL ...]   GPU tests, then one bench run per argument
# (each argument is an env assignment applied to that run; "-" = defaults)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
show() { python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
k=d['kernel_ms_per_step']
print(sys.argv[1], round(d['value'],1), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'])
print('  ', {n:v for n,v in list(k.items())[:16]})
PY
}
i=0
for cfg in "${@:--}"; do
  i=$((i+1))
  if [ "$cfg" = "-" ]; then e=""; else e="$cfg"; fi
  env $e timeout 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_ab$i.json 2>gpurun_out/bench_ab$i.err
  echo "[$cfg]"; show gpurun_out/bench_ab$i.json; tail -2 gpurun_out/bench_ab$i.err
done
