#!/bin/bash
for tag in "" NO_TMA NO_MMA; do
  DINOX_LIB_TAG=$tag DINOX_PAIR=0 timeout 120 python tools/probe_time.py 2>&1 | tail -1
done
