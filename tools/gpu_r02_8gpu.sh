#!/bin/bash
# round 2, 8 GPUs: sharded == full batch over NCCL, the bench line with the C3 / C4 extras, the all-reduce placement
# A/B, and the fused dW2 GEMM + reduce-scatter window against the NCCL path
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29701 tools/dist_parity_nccl.py > gpurun_out/r02_dist_parity_${N}gpu.txt 2> gpurun_out/dp.err; tail -3 gpurun_out/r02_dist_parity_${N}gpu.txt
timeout 900 $TR --master-port 29702 bench.py --gpus $N --steps 60 --warmup 4 --no-hbm-table > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/b8.err; echo bench rc=$?
DINOX_CENTER_AR=late timeout 900 $TR --master-port 29703 bench.py --gpus $N --steps 60 --warmup 4 --no-hbm-table --no-extra --sustained-s 0 > gpurun_out/r02_bench_${N}gpu_ar_late.json 2> gpurun_out/b8l.err; echo bench-late rc=$?
timeout 600 $TR --master-port 29704 tools/dist_fused_rs.py time 2> gpurun_out/rs.err | grep -v "^\*\|OMP_NUM" | tail -5 > gpurun_out/r02_fused_rs_${N}gpu.txt; cat gpurun_out/r02_fused_rs_${N}gpu.txt
python - <<PY
import json
for f in ("gpurun_out/r02_bench_${N}gpu.json", "gpurun_out/r02_bench_${N}gpu_ar_late.json"):
    try:
        d = json.load(open(f))
        print(f, round(d["ms_per_step"], 4), round(d["value"]), "e2e", round(d["e2e"]["value"]), d["per_rank_ms"], d.get("extra"))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 gpurun_out/b8.err
