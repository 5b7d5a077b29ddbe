#!/bin/bash
mkdir -p gpurun_out
for tag in "" epi8 epi16 mmaepi8 tmaepi8 full16; do
  DINOX_LIB_TAG=$tag DINOX_PAIR=1 timeout 120 python tools/probe_time.py 2>&1 | tail -1
done | tee gpurun_out/probe_time_parts2.log
