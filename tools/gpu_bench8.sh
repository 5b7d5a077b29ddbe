#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 60 --warmup 4 --no-cpu-baseline > gpurun_out/final_bench_${N}gpu.json 2> gpurun_out/final_bench_${N}gpu.err; echo rc=$?
python - $N <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/final_bench_{sys.argv[1]}gpu.json")); print(round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["clocks"])
PY
