#!/bin/bash
mkdir -p gpurun_out
for pair in 0 1; do
  echo "=== DINOX_PAIR=$pair"
  DINOX_PAIR=$pair bash tools/run_probes.sh gemm stats grad > gpurun_out/probes_$pair.log 2>&1
  grep -E "MISMATCH|EXC|rror|timeout|time |exit|stats rows|grad E" gpurun_out/probes_$pair.log | head -40
done
DINOX_PAIR=0 bash tools/gpu_quick.sh
