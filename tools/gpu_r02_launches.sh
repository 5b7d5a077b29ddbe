#!/bin/bash
# launch list of the bench command (per-launch device times; cold-cache, serialised: compare SHARES)
mkdir -p gpurun_out
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --sustained-s 0 --no-hbm-table > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --sustained-s 0 --no-hbm-table > gpurun_out/ncu_launches.log 2>&1
tail -c 300 gpurun_out/plain.log; wc -l gpurun_out/r02_launches.csv
