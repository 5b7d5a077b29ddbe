#!/bin/bash
mkdir -p gpurun_out
for d in 256 384 128; do for r in 0 6 0 6; do
  PROBE_D=$d DINOX_RESA=$r timeout 200 python tools/probe_grad.py 2>&1 | tail -1
done; done | tee gpurun_out/resb_dsweep.log
