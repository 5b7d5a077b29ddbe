#!/bin/bash
# round 2: plain timings, then one full ncu capture of the read-back pass pair + backward GEMMs
mkdir -p gpurun_out
python tools/probe_r02.py time > gpurun_out/r02_probe_time.log 2>&1; cat gpurun_out/r02_probe_time.log | tail -2
python tools/probe_r02.py once > gpurun_out/r02_probe_once.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel --launch-skip 2 -c 6 -f -o gpurun_out/r02_prof \
    python tools/probe_r02.py once > gpurun_out/r02_prof_ncu.log 2>&1
tail -2 gpurun_out/r02_prof_ncu.log; ls -la gpurun_out/r02_prof.ncu-rep
