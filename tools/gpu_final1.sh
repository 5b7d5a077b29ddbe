#!/bin/bash
# final single-GPU pass of a round: GPU tests, smoke, default bench line (with the CPU baseline leg), launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/final_bench_1gpu.json 2> gpurun_out/final_bench_1gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/final_bench_1gpu.json'))
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'frac',round(d['roofline']['frac'],3),'step_frac',round(d['roofline']['step_frac'],3),'cpu',d['cpu_baseline'] and round(d['cpu_baseline']['value'],1),d['clocks'])
print({k:round(v,3) for k,v in d['roofline']['kernels_ms'].items()}, 'ema_gbs', round(d['roofline']['ema_gbs']))
PY
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
wc -l gpurun_out/launches.csv
