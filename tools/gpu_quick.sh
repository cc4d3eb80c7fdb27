mkdir -p gpurun_out; timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2; timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_x.json; python -c "
import json; d=json.load(open('gpurun_out/bench_x.json')); print(d['value'], d['e2e']['value'], d['clocks']); print({k:v for k,v in list(d['kernel_ms_per_step'].items())[:12]})"
