#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_gpu57.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu57.log
SWEEP_PAIRS=1 timeout 600 python tools/k2_sweep.py cfg2 'PLS_K4_DIRECT=0' '' 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('pairs cfg2', d['setting'], round(d['ms'],3), d.get('same_winner'), d.get('pivots'))"
timeout 600 python tools/k2_sweep.py cfg2 'PLS_K4_DIRECT=0' '' 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('literal cfg2', d['setting'], round(d['ms'],3), d.get('same_winner'))"
SWEEP_COUNT_LOG2=18 timeout 600 python tools/k2_sweep.py m512k24 'PLS_K4_DIRECT=0' '' 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('m512k24 2^18', d['setting'], round(d['ms'],3), d.get('same_winner'))"
timeout 1200 python tools/v4_fuzz.py 60 99 > gpurun_out/v4_fuzz5.jsonl 2> gpurun_out/v4_fuzz5.err; echo "fuzz rc=$?"; tail -1 gpurun_out/v4_fuzz5.jsonl | cut -c1-200; tail -2 gpurun_out/v4_fuzz5.err
