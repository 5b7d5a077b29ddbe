#!/bin/bash
for tag in "" fld fmath; do for pair in 0; do
  DINOX_LIB_TAG=$tag DINOX_PAIR=$pair timeout 120 python tools/probe_time.py 2>&1 | tail -1
done; done | tee gpurun_out/probe_time_epiparts.log
