#!/bin/bash
# Round profile capture: (1) launch list of the bench command, (2) full ncu sections of the five big GEMM launches.
mkdir -p gpurun_out
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python tools/probe_prof.py 1 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 5 -o gpurun_out/prof_round \
    python tools/probe_prof.py 1 > gpurun_out/prof_ncu.log 2>&1
tail -c 200 gpurun_out/plain.log; tail -2 gpurun_out/prof_ncu.log; wc -l gpurun_out/launches.csv
