#!/bin/bash
for pair in 1 7; do DINOX_PAIR=$pair timeout 120 python tools/probe_time.py 2>&1 | tail -1; done
python tools/probe_pairdbg.py 2>&1 | grep -v OK | head
for i in 1 2; do DINOX_PAIR=7 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -2; done
