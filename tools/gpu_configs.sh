#!/bin/bash
# bench lines of the other BASELINE.json configs on one GPU (per-GPU shapes of C3 / C4, low end of the C5 sweep)
mkdir -p gpurun_out
run() { name=$1; shift; timeout 500 python bench.py --steps 40 --warmup 4 --no-cpu-baseline "$@" > gpurun_out/cfg_$name.json 2> gpurun_out/cfg_$name.err; echo "rc=$? $name"
python - $name <<'PY'
import json,sys
d=json.load(open(f'gpurun_out/cfg_{sys.argv[1]}.json'))
print(' value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'frac',round(d['roofline']['frac'],3),{k:round(v,3) for k,v in d['roofline']['kernels_ms'].items()})
PY
}
run C2
run C3 --config C3 --teacher-mode sinkhorn
run C4 --config C4
run C5lo --config C5lo
