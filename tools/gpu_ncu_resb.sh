#!/bin/bash
mkdir -p gpurun_out
DINOX_RESA=6 timeout 200 python tools/probe_grad.py 2>&1 | tail -1 && \
DINOX_RESA=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 9 -c 1 -o gpurun_out/prof_resb python tools/probe_grad.py > gpurun_out/prof_resb.log 2>&1
tail -2 gpurun_out/prof_resb.log
