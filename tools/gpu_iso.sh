#!/bin/bash
mkdir -p gpurun_out
DINOX_LIB_TAG=gef timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --no-header -p no:cacheprovider -x -k "fused" 2>&1 | tail -2
for tag in default gef default gef; do
  if [ $tag = default ]; then timeout 200 python tools/probe_time.py 2>&1 | tail -1; else DINOX_LIB_TAG=$tag timeout 200 python tools/probe_time.py 2>&1 | tail -1; fi
done | tee gpurun_out/iso_evict_first.log
