#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
python bench.py --steps 12 --warmup 4 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo "bench exit $?" >> gpurun_out/bench_r1.err
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 8 -c 8 -o gpurun_out/prof_gemm \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/smoke.log; cat gpurun_out/bench_r1.json; tail -3 gpurun_out/bench_r1.err; tail -2 gpurun_out/ncu1.log; tail -2 gpurun_out/ncu2.log
