#!/bin/bash
mkdir -p gpurun_out
for pair in 1; do
  echo "=== DINOX_PAIR=$pair"
  DINOX_PAIR=$pair bash tools/run_probes.sh gemm stats grad > gpurun_out/probes_$pair.log 2>&1
  grep -E "MISMATCH|EXC|Error|error|timeout|time |exit|dW2|dH" gpurun_out/probes_$pair.log | head -40
done
python tools/probe_prof.py 1 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 5 -o gpurun_out/prof_r1b \
    python tools/probe_prof.py 1 > gpurun_out/prof_ncu.log 2>&1
tail -2 gpurun_out/prof_plain.log; tail -2 gpurun_out/prof_ncu.log
