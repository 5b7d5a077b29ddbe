#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --no-header -p no:cacheprovider -x -k "sharded or adamw" 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 tools/dist_sharded_adamw.py 2>&1 | grep -v "^W\|^\*\|OMP_NUM\|^$" | tee gpurun_out/dist_sharded_adamw.log | tail -22
