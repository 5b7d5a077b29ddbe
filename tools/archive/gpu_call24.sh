#!/bin/bash
for tag in ld math; do
  DINOX_LIB_TAG=$tag DINOX_PAIR=1 timeout 120 python tools/probe_time.py 2>&1 | tail -1
done | tee gpurun_out/probe_time_parts3.log
