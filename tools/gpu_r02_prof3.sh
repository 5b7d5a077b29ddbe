#!/bin/bash
mkdir -p gpurun_out
DINOX_RB_SCHED=3 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"EpiGradR" --launch-skip 0 -c 1 -f -o gpurun_out/r02_prof_g2 \
    python tools/probe_r02.py once > gpurun_out/r02_prof_ncu.log 2>&1
tail -1 gpurun_out/r02_prof_ncu.log; ls -la gpurun_out/r02_prof_g2.ncu-rep
