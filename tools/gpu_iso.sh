#!/bin/bash
mkdir -p gpurun_out
for tag in default noepi nomma notma; do for r in 0 6; do
  if [ $tag = default ]; then DINOX_RESA=$r timeout 200 python tools/probe_grad.py 2>&1 | tail -1; else DINOX_LIB_TAG=$tag DINOX_RESA=$r timeout 200 python tools/probe_grad.py 2>&1 | tail -1; fi
done; done | tee gpurun_out/iso_resb.log
