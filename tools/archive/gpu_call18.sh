#!/bin/bash
mkdir -p gpurun_out
for tag in only_mma only_alt1 only_alt2; do
  DINOX_LIB_TAG=$tag DINOX_PAIR=0 timeout 120 python tools/probe_time.py 2>&1 | tail -1
done | tee gpurun_out/probe_time_alt.log
