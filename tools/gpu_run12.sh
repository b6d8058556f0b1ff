#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
V3="PLS_K2_IMPL=v3,PLS_K3_QS=0"
timeout 900 python tools/k2_sweep.py cfg2 "" "$V3,PLS_K3_T=128,PLS_K3_MINB=4" "$V3,PLS_K3_T=128,PLS_K3_MINB=6" "$V3,PLS_K3_T=256,PLS_K3_MINB=3" "$V3,PLS_K3_T=256,PLS_K3_MINB=2" "PLS_K2_IMPL=v3,PLS_K3_QS=60,PLS_K3_T=256,PLS_K3_MINB=3" "PLS_K2_IMPL=v2" 2> gpurun_out/sweep12.err | tee gpurun_out/k2_sweep12_cfg2.jsonl
timeout 900 python tools/k2_sweep.py m512k16 "" "PLS_K3_QS=0,PLS_K3_MINB=3" 2>> gpurun_out/sweep12.err | tee gpurun_out/k2_sweep12_m512.jsonl
timeout 900 python tools/k2_sweep.py cfg2 "PLS_K2_PHASES=1" 2> gpurun_out/phases_v3b_cfg2.txt | tail -1
tail -3 gpurun_out/sweep12.err
