#!/bin/bash
# A/B of the side-stream concurrency levels (DINOX_CONCURRENCY) on one box
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x 2>&1 | tail -8 > gpurun_out/conc_pytest.log; tail -2 gpurun_out/conc_pytest.log
for c in 0 1 2 0 1 2; do
  DINOX_CONCURRENCY=$c timeout 400 python bench.py --steps 60 --warmup 4 --no-cpu-baseline > gpurun_out/conc_bench_$c.json 2> gpurun_out/conc_bench_$c.err
  python - $c <<'PY'
import json,sys
d=json.load(open(f'gpurun_out/conc_bench_{sys.argv[1]}.json'))
print('CONC',sys.argv[1],'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'clk',d['clocks']['sm_mhz'])
PY
done
