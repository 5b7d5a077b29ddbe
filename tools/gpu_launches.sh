#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -c 300 gpurun_out/plain.log; wc -l gpurun_out/launches.csv
