#!/bin/bash
# A/B of the resident-A mode (DINOX_RESA bit mask) on one box: parity of the two loss-head passes, then timings
mkdir -p gpurun_out
for r in 2 6; do
  DINOX_RESA=$r timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q --no-header -rf -p no:cacheprovider -x 2>&1 | tail -6 > gpurun_out/resa_pytest_$r.log
  echo "RESA=$r: $(tail -1 gpurun_out/resa_pytest_$r.log)"
done
for r in 0 2 6 0 2 6; do
  DINOX_RESA=$r timeout 300 python tools/probe_time.py 2>&1 | tail -1 | sed "s/^/RESA=$r /" | tee -a gpurun_out/resa_time.log
done
