#!/bin/bash
# step-level A/B of one environment knob: tools/gpu_ab_env.sh VAR val1 val2 [reps]
mkdir -p gpurun_out
VAR=$1; A=$2; B=$3; R=${4:-3}
for i in $(seq $R); do for v in $A $B; do
  env $VAR=$v timeout 400 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - $VAR $v <<'PY'
import json,sys
d=json.load(open('gpurun_out/ab.json'))
print(sys.argv[1],sys.argv[2],'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'grad',round(d['roofline']['kernels_ms']['head_grad'],3))
PY
done; done | tee gpurun_out/ab_$VAR.log
