#!/bin/bash
# run every probe group in its own process with a timeout; logs under gpurun_out/
mkdir -p gpurun_out
for g in "$@"; do
  echo "=== $g ===" | tee gpurun_out/probe_$g.log
  timeout 300 python tools/probe.py $g >> gpurun_out/probe_$g.log 2>&1
  echo "exit $?" >> gpurun_out/probe_$g.log
  tail -n 60 gpurun_out/probe_$g.log
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> gpurun_out/probe_env.log 2>&1
