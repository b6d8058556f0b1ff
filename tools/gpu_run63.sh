#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu63.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu63.log
SWEEP_PAIRS=1 timeout 300 python tools/k2_sweep.py cfg2 '' 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('pairs cfg2', round(d['ms'],3), d.get('same_winner'))"
SWEEP_PAIRS=1 SWEEP_COUNT_LOG2=21 timeout 300 python tools/k2_sweep.py k20 '' 'PLS_K2_IMPL=v3' 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('pairs k20', d['setting'], round(d['ms'],3), d.get('same_winner'))"
SWEEP_PAIRS=1 SWEEP_COUNT_LOG2=20 timeout 300 python tools/k2_sweep.py m512k24 '' 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('pairs m512k24 2^19 problems', round(d['ms'],3), d.get('same_winner'))"
timeout 600 python tools/v4_fuzz.py 60 5 > gpurun_out/v4_fuzz6.jsonl 2> gpurun_out/v4_fuzz6.err; echo "fuzz rc=$?"; tail -1 gpurun_out/v4_fuzz6.jsonl | cut -c1-160
