#!/bin/bash
mkdir -p gpurun_out
for L2 in 14 15 16; do
SWEEP_COUNT_LOG2=$L2 timeout 300 python tools/k2_sweep.py cfg2 'PLS_K2_IMPL=v3' 'PLS_K2_IMPL=v4' 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print($L2, d['setting'], round(d['ms'],3), d.get('same_winner'))"
done
for L2 in 14 15 16; do
SWEEP_COUNT_LOG2=$L2 timeout 300 python tools/k2_sweep.py m512k16 'PLS_K2_IMPL=v3' 'PLS_K2_IMPL=v4' 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('m512', $L2, d['setting'], round(d['ms'],3), d.get('same_winner'))"
done
