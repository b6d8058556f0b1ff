#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
V3="PLS_K2_IMPL=v3,PLS_K3_QS=0"
timeout 900 python tools/k2_sweep.py cfg2 "" "$V3,PLS_K3_T=128,PLS_K3_MINB=4" "$V3,PLS_K3_T=256,PLS_K3_MINB=3" "$V3,PLS_K3_T=256,PLS_K3_MINB=2" "PLS_K2_IMPL=v3,PLS_K3_QS=60,PLS_K3_T=256,PLS_K3_MINB=3" 2> gpurun_out/sweep14.err | tee gpurun_out/k2_sweep14_cfg2.jsonl
timeout 900 python tools/k2_sweep.py m512k16 "" "PLS_K3_QS=0,PLS_K3_MINB=3" 2>> gpurun_out/sweep14.err | tee gpurun_out/k2_sweep14_m512.jsonl
timeout 900 python tools/k2_sweep.py cfg2 "PLS_K2_PHASES=1" 2> gpurun_out/phases_v3d_cfg2.txt | tail -1
