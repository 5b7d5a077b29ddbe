#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x 2>&1 | tail -5
for i in 1 2; do for ps in rows tokens; do
  timeout 400 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --patch-source $ps > gpurun_out/ps_$ps.json 2> gpurun_out/ps_$ps.err
  python - $ps <<'PY'
import json,sys
d=json.load(open(f'gpurun_out/ps_{sys.argv[1]}.json'))
print(sys.argv[1],'value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'h2d',d['e2e']['h2d_bytes_per_step'],'launches',d['gpu_launches'])
PY
done; done
