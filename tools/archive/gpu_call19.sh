#!/bin/bash
mkdir -p gpurun_out
for tag in "" ew8 mma; do for pair in 1 7; do
  DINOX_LIB_TAG=$tag DINOX_PAIR=$pair timeout 120 python tools/probe_time.py 2>&1 | tail -1
done; done | tee gpurun_out/probe_time.log
bash tools/run_probes.sh gemm stats grad > gpurun_out/probes.log 2>&1
grep -E "MISMATCH|EXC|Error|error|timeout|exit" gpurun_out/probes.log | head -20
bash tools/gpu_quick.sh
