#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "v4 or twolevel or toy" > gpurun_out/pytest_gpu37.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu37.log
timeout 600 python tools/k2_sweep.py cfg2 '' 'PLS_K4_L=6' > gpurun_out/k2_sweep37_cfg2.jsonl 2> gpurun_out/sweep37.err
cut -c1-200 gpurun_out/k2_sweep37_cfg2.jsonl
SWEEP_COUNT_LOG2=21 timeout 600 python tools/k2_sweep.py k20 '' 'PLS_K4_L=6' > gpurun_out/k2_sweep37_k20.jsonl 2>> gpurun_out/sweep37.err
cut -c1-200 gpurun_out/k2_sweep37_k20.jsonl
SWEEP_COUNT_LOG2=20 timeout 600 python tools/k2_sweep.py m512k24 '' > gpurun_out/k2_sweep37_m512k24.jsonl 2>> gpurun_out/sweep37.err
cut -c1-200 gpurun_out/k2_sweep37_m512k24.jsonl
