#!/bin/bash
mkdir -p gpurun_out
export DINOX_PAIR=1
python tools/probe_prof.py 1 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 3 -o gpurun_out/prof_r1b_pair \
    python tools/probe_prof.py 1 > gpurun_out/prof_ncu.log 2>&1
tail -2 gpurun_out/prof_plain.log; tail -2 gpurun_out/prof_ncu.log
