#!/bin/bash
for c in 0 2; do DINOX_BENCH_DEBUG=1 DINOX_CONCURRENCY=$c timeout 300 python bench.py --steps 20 --warmup 4 --no-cpu-baseline 2>&1 >/dev/null | grep timer; done
