#!/bin/bash
mkdir -p gpurun_out
SWEEP_PAIRS=1 timeout 600 python tools/k2_sweep.py cfg2 '' 'PLS_K4_OCC=3' 'PLS_K4_L=4' 'PLS_K4_L=6' 'PLS_K4_L=7' 'PLS_K4_L=8' 'PLS_K2_IMPL=v3' 'PLS_K4_L=6,PLS_K2_PHASES=1' 2> gpurun_out/sweep47.err | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['setting'], round(d['ms'],3), d.get('same_winner'), d.get('pivots'))"
tail -24 gpurun_out/sweep47.err
