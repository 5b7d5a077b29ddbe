#!/bin/bash
mkdir -p gpurun_out
bash tools/run_probes.sh gemm stats grad > gpurun_out/probes.log 2>&1
grep -E "MISMATCH|EXC|Error|error|timeout|time |exit" gpurun_out/probes.log | head -40
bash tools/gpu_quick.sh
