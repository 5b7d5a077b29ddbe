#!/bin/bash
# A/B of the A-operand collector hint (NSPLIT=3 tiles): parity, then timings vs a build without it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_fullsize.py -m gpu -q --no-header -rf -p no:cacheprovider -x 2>&1 | tail -4
for t in nocoll default nocoll default; do
  if [ $t = default ]; then timeout 300 python tools/probe_time.py 2>&1 | tail -1; else DINOX_LIB_TAG=$t timeout 300 python tools/probe_time.py 2>&1 | tail -1; fi
done | tee gpurun_out/coll_time.log
