#!/bin/bash
mkdir -p gpurun_out
DINOX_PDL=1 timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x 2>&1 | tail -3
bash tools/gpu_ab_env.sh DINOX_PDL 0 1 3
