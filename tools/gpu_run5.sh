#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python tools/k2_sweep.py cfg2 'PLS_K2_IMPL=v2' 'PLS_K2_IMPL=v3,PLS_K3_QS=0' 'PLS_K2_IMPL=v3,PLS_K3_QS=0,PLS_K3_MINB=3' 'PLS_K2_IMPL=v3,PLS_K3_QS=0,PLS_K3_T=512' 'PLS_K2_IMPL=v3,PLS_K3_QS=91' 'PLS_K2_IMPL=v3,PLS_K3_QS=120' 'PLS_K2_IMPL=v3,PLS_K3_QS=-1,PLS_K3_T=512' 2>&1 | tee gpurun_out/k2_sweep_cfg2.jsonl
timeout 900 python tools/k2_sweep.py m512 'PLS_K3_QS=0' 'PLS_K3_QS=0,PLS_K3_T=512' 'PLS_K3_QS=100' 'PLS_K3_QS=-1,PLS_K3_T=512' 2>&1 | tee gpurun_out/k2_sweep_m512.jsonl
