#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "v4 or toy" > gpurun_out/pytest_gpu30.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu30.log
timeout 600 python tools/k2_sweep.py cfg2 '' 'PLS_K4_L=6' > gpurun_out/k2_sweep30_cfg2.jsonl 2> gpurun_out/sweep30.err
cut -c1-200 gpurun_out/k2_sweep30_cfg2.jsonl
SWEEP_COUNT_LOG2=20 timeout 900 python tools/k2_sweep.py m512k24 '' 'PLS_K4_T=512' 'PLS_K4_T=512,PLS_K4_L=7' 'PLS_K4_OCC=1' > gpurun_out/k2_sweep30_m512k24.jsonl 2>> gpurun_out/sweep30.err; echo "sweep rc=$?"
cut -c1-220 gpurun_out/k2_sweep30_m512k24.jsonl
SWEEP_COUNT_LOG2=18 ncu --set full --clock-control none --import-source on -k regex:k2v4_orthant -c 1 -f -o gpurun_out/k2_prof30_m512k24 python tools/k2_sweep.py m512k24 '' > gpurun_out/ncu30.log 2>&1
tail -2 gpurun_out/ncu30.log
