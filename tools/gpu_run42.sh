#!/bin/bash
mkdir -p gpurun_out
SWEEP_COUNT_LOG2=21 timeout 600 python tools/k2_sweep.py k20 'PLS_K4_OCC=1' 'PLS_K4_OCC=2' 'PLS_K4_OCC=3' '' 'PLS_K4_OCC=2,PLS_K4_QS=36' 'PLS_K4_OCC=2,PLS_K4_T=256' 'PLS_K4_T=256,PLS_K4_MINB=2,PLS_K4_QS=36' 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['setting'], round(d['ms'],2), d.get('same_winner'))"
