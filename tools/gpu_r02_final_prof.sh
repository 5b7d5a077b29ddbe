#!/bin/bash
# round-2 profile set of the final code: plain bench, ncu launch list of the same command, full ncu of the big
# kernels (tools/probe_r02.py once), dram bytes / duration of the HBM-bound reductions (tools/probe_hbm.py),
# in-graph timeline of the micro-step (tools/trace_step.py)
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/bench.err; echo bench rc=$?
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --sustained-s 0 --no-hbm-table --no-extra --no-torch-eager > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --sustained-s 0 --no-hbm-table --no-extra --no-torch-eager > gpurun_out/ncu_launches.log 2>&1
python tools/trace_step.py --isolated --out gpurun_out/r02_step_timeline.txt > gpurun_out/trace_iso.log 2>&1; tail -1 gpurun_out/r02_step_timeline.txt
python tools/trace_step.py --out gpurun_out/r02_step_timeline_steady.txt > gpurun_out/trace_steady.log 2>&1; tail -1 gpurun_out/r02_step_timeline_steady.txt
python tools/probe_r02.py time > gpurun_out/r02_probe_time.log 2>&1; tail -1 gpurun_out/r02_probe_time.log
python tools/probe_r02.py once > gpurun_out/r02_probe_once.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel --launch-skip 2 -c 5 -f -o gpurun_out/r02_prof \
    python tools/probe_r02.py once > gpurun_out/r02_prof_ncu.log 2>&1
python tools/probe_hbm.py > gpurun_out/hbm_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_hbm_kernels_ncu.csv \
    python tools/probe_hbm.py > gpurun_out/hbm_ncu.log 2>&1
ls -la gpurun_out/r02_prof.ncu-rep gpurun_out/r02_launches.csv gpurun_out/r02_hbm_kernels_ncu.csv
