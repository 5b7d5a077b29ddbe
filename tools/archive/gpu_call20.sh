#!/bin/bash
mkdir -p gpurun_out
python tools/probe_pairdbg.py 2>&1 | tail -12
for pair in 0 1 3 7; do
  DINOX_PAIR=$pair timeout 120 python tools/probe_time.py 2>&1 | tail -1
done | tee gpurun_out/probe_time.log
DINOX_PAIR=7 bash tools/run_probes.sh gemm stats grad > gpurun_out/probes.log 2>&1
grep -E "MISMATCH|EXC|Error|error|timeout|exit" gpurun_out/probes.log | head -20
DINOX_PAIR=7 python -m pytest tests -m gpu -q --no-header -x -p no:cacheprovider 2>&1 | tail -3
bash tools/gpu_quick.sh
