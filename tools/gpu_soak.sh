#!/bin/bash
# flakiness soak: the GPU suite several times (random order seeds differ only by run), the bench loss must repeat bit for bit
mkdir -p gpurun_out
for i in 1 2 3 4 5; do timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -1; done
for i in 1 2 3; do timeout 300 python bench.py --steps 40 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('loss', repr(d['loss']), 'ms', round(d['ms_per_step'],4))"; done
