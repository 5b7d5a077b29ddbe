#!/bin/bash
mkdir -p gpurun_out
python tools/probe_prof.py 2 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 5 -c 5 -o gpurun_out/prof_big \
    python tools/probe_prof.py 2 > gpurun_out/prof_ncu.log 2>&1
tail -3 gpurun_out/prof_plain.log; tail -3 gpurun_out/prof_ncu.log
