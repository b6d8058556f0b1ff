#!/bin/bash
mkdir -p gpurun_out
V3="PLS_K2_IMPL=v3,PLS_K3_QS=0"
timeout 900 python tools/k2_sweep.py cfg2 "" "$V3,PLS_K3_T=128,PLS_K3_MINB=4" "$V3,PLS_K3_T=256,PLS_K3_MINB=3" "$V3,PLS_K3_T=256,PLS_K3_MINB=2" 2> gpurun_out/sweep13.err | tee gpurun_out/k2_sweep13_cfg2.jsonl
timeout 900 python tools/k2_sweep.py m512k16 "" 2>> gpurun_out/sweep13.err | tee gpurun_out/k2_sweep13_m512.jsonl
timeout 900 python tools/k2_sweep.py cfg2 "$V3,PLS_K3_T=128,PLS_K3_MINB=4,PLS_K2_PHASES=1" 2> gpurun_out/phases_v3c_cfg2.txt | tail -1
