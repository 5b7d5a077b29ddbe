#!/bin/bash
# A/B of the side-stream concurrency levels (DINOX_CONCURRENCY) on one box: GPU suite at the non-default level, then timings
mkdir -p gpurun_out
L=${1:-3}
DINOX_CONCURRENCY=$L timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x 2>&1 | tail -3
bash tools/gpu_ab_env.sh DINOX_CONCURRENCY 2 $L 3
