#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu20.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu20.log
timeout 600 python tools/k2_sweep.py cfg2 'PLS_K3_PROW16=0' '' 'PLS_K3_PROW16=104' 'PLS_K3_PROW16=-1' 'PLS_K3_PROW16=0,PLS_K2_PHASES=1' 'PLS_K2_PHASES=1' > gpurun_out/k2_sweep20_cfg2.jsonl 2> gpurun_out/sweep20.err; echo "sweep rc=$?"
cut -c1-330 gpurun_out/k2_sweep20_cfg2.jsonl
timeout 600 python tools/k2_sweep.py m512k16 'PLS_K3_PROW16=0' '' 'PLS_K3_PROW16=-1' > gpurun_out/k2_sweep20_m512.jsonl 2>> gpurun_out/sweep20.err; echo "sweep rc=$?"
cut -c1-330 gpurun_out/k2_sweep20_m512.jsonl
tail -3 gpurun_out/sweep20.err
