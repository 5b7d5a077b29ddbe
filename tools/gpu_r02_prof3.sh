#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"EpiGradR" --launch-skip 1 -c 1 -f -o gpurun_out/r02_prof_g2 \
    python tools/probe_r02.py once > gpurun_out/r02_prof_ncu.log 2>&1
tail -1 gpurun_out/r02_prof_ncu.log; ls -la gpurun_out/r02_prof_g2.ncu-rep
