#!/bin/bash
mkdir -p gpurun_out
python tools/probe_pairdbg.py 2>&1 | tail -12
for pair in 0 7; do DINOX_PAIR=$pair python -m pytest tests -m gpu -q --no-header -x -p no:cacheprovider 2>&1 | tail -3; done
for pair in 1 7; do
  DINOX_PAIR=$pair timeout 120 python tools/probe_time.py 2>&1 | tail -1
done | tee gpurun_out/probe_time.log
bash tools/gpu_quick.sh
