#!/bin/bash
mkdir -p gpurun_out
PLS_K4_T=64 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "twolevel" > gpurun_out/pytest_gpu36.log 2>&1; echo "pytest(T=64) rc=$?"; tail -3 gpurun_out/pytest_gpu36.log
timeout 600 python tools/k2_sweep.py cfg2 '' 'PLS_K4_T=64' 'PLS_K4_T=64,PLS_K4_MINB=4' 'PLS_K4_T=64,PLS_K4_MINB=3' 'PLS_K4_T=64,PLS_K4_L=6' > gpurun_out/k2_sweep36_cfg2.jsonl 2> gpurun_out/sweep36.err
cut -c1-200 gpurun_out/k2_sweep36_cfg2.jsonl
SWEEP_COUNT_LOG2=21 timeout 600 python tools/k2_sweep.py k20 '' 'PLS_K4_T=64' 'PLS_K4_T=64,PLS_K4_L=6' > gpurun_out/k2_sweep36_k20.jsonl 2>> gpurun_out/sweep36.err
cut -c1-200 gpurun_out/k2_sweep36_k20.jsonl
