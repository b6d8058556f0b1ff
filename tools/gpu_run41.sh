#!/bin/bash
mkdir -p gpurun_out
export PLS_K4_DRIFT=1
for v in 128 1000000; do
PLS_K4_VERIFY=$v timeout 300 python tools/k2_sweep.py cfg2 "PLS_K4_VERIFY=$v" 2>&1 | grep -E "max KKT|same_winner" | tail -2 | cut -c1-200
PLS_K4_VERIFY=$v SWEEP_COUNT_LOG2=21 timeout 300 python tools/k2_sweep.py k20 "PLS_K4_VERIFY=$v" 2>&1 | grep -E "max KKT|same_winner" | tail -2 | cut -c1-200
PLS_K4_VERIFY=$v SWEEP_COUNT_LOG2=20 timeout 300 python tools/k2_sweep.py m512k24 "PLS_K4_VERIFY=$v" 2>&1 | grep -E "max KKT|same_winner" | tail -2 | cut -c1-200
done
