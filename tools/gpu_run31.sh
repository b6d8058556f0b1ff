#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "v4 or toy" > gpurun_out/pytest_gpu31.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu31.log
timeout 600 python tools/k2_sweep.py cfg2 '' 'PLS_K4_L=4' 'PLS_K4_L=6' > gpurun_out/k2_sweep31_cfg2.jsonl 2> gpurun_out/sweep31.err
cut -c1-200 gpurun_out/k2_sweep31_cfg2.jsonl
SWEEP_COUNT_LOG2=20 timeout 900 python tools/k2_sweep.py m512k24 '' 'PLS_K4_L=5' 'PLS_K4_L=4' > gpurun_out/k2_sweep31_m512k24.jsonl 2>> gpurun_out/sweep31.err; echo "sweep rc=$?"
cut -c1-220 gpurun_out/k2_sweep31_m512k24.jsonl
