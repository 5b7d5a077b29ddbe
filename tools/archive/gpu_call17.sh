#!/bin/bash
mkdir -p gpurun_out
for tag in "" ew8; do
  DINOX_LIB_TAG=$tag timeout 120 python tools/probe_time.py 2>&1 | tail -1
done | tee gpurun_out/probe_time.log
bash tools/run_probes.sh stats grad > gpurun_out/probes.log 2>&1
grep -E "MISMATCH|EXC|Error|error|timeout|exit|time" gpurun_out/probes.log | head -20
bash tools/gpu_quick.sh
