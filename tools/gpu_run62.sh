#!/bin/bash
SWEEP_PAIRS=1 timeout 600 python tools/k2_sweep.py cfg2 '' 'PLS_K4_T=512' 'PLS_K4_T=256,PLS_K4_MINB=3,PLS_K4_L=4' 'PLS_K4_T=256,PLS_K4_MINB=3,PLS_K4_L=6' 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('pairs cfg2', d['setting'], round(d['ms'],3), d.get('same_winner'))"
