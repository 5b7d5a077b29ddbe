#!/bin/bash
# GPU suite + two default-length bench runs (no CPU leg)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x 2>&1 | tail -4
for i in 1 2; do
  timeout 400 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/quick.json 2> gpurun_out/quick.err
  python - <<'PY'
import json
d=json.load(open('gpurun_out/quick.json'))
print('value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),{k:round(v,3) for k,v in d['roofline']['kernels_ms'].items()},'launches',d['gpu_launches'])
PY
done
DINOX_RESA=0 timeout 400 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/quick0.json 2> gpurun_out/quick0.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/quick0.json'))
print('RESA=0 value',round(d['value']),'ms',round(d['ms_per_step'],4))
PY
