#!/bin/bash
# two-GPU checks: NCCL data-parallel parity, then the bench in centre and Sinkhorn mode
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_parity_nccl.py 2>&1 | grep -i "pass\|fail\|error" | tail -12
for mode in center sinkhorn; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 40 --warmup 4 --no-cpu-baseline --teacher-mode $mode > gpurun_out/dist2_$mode.json 2> gpurun_out/dist2_$mode.err
  echo "rc=$? $mode"; tail -c 900 gpurun_out/dist2_$mode.json | head -c 700; echo
done
