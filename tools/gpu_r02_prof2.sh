#!/bin/bash
# full ncu capture of the read-back pair only (both schedules of pass 2)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"EpiGradR|EpiTeachQ" --launch-skip 2 -c 2 -f -o gpurun_out/r02_prof_rb \
    python tools/probe_r02.py once > gpurun_out/r02_prof_ncu.log 2>&1
DINOX_RB_SCHED=3 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"EpiGradR" --launch-skip 1 -c 1 -f -o gpurun_out/r02_prof_rb_cols \
    python tools/probe_r02.py once >> gpurun_out/r02_prof_ncu.log 2>&1
tail -2 gpurun_out/r02_prof_ncu.log; ls -la gpurun_out/*.ncu-rep
