#!/bin/bash
mkdir -p gpurun_out
for tag in "" nofwd; do for pair in 0 1; do
  DINOX_LIB_TAG=$tag DINOX_PAIR=$pair timeout 120 python tools/probe_time.py 2>&1 | tail -1
done; done | tee gpurun_out/probe_time.log
DINOX_PAIR=1 python -m pytest tests -m gpu -q --no-header -x -p no:cacheprovider 2>&1 | tail -3
