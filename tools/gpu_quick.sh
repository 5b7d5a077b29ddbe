#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
python bench.py --steps 40 --warmup 4 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -2 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'clk',d['clocks'])
print({k:round(v,4) for k,v in d['roofline']['kernels_ms'].items()})
print('launches',d['gpu_launches'],'frac',d['roofline']['frac'],'step_frac',d['roofline']['step_frac'])
PY
