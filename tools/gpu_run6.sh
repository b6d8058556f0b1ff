#!/bin/bash
mkdir -p gpurun_out
V3="PLS_K2_IMPL=v3,PLS_K3_QS=0"
timeout 900 python tools/k2_sweep.py cfg2 "$V3,PLS_K3_MINB=3" "$V3,PLS_K3_MINB=3,PLS_K3_CHAIN=6" "$V3,PLS_K3_MINB=3,PLS_K3_CHAIN=7" "$V3,PLS_K3_MINB=4" "$V3,PLS_K3_T=128,PLS_K3_MINB=4" "$V3,PLS_K3_T=128,PLS_K3_MINB=5" "$V3,PLS_K3_T=128,PLS_K3_MINB=6" "PLS_K2_IMPL=v3,PLS_K3_QS=60,PLS_K3_MINB=3" "PLS_K2_IMPL=v3,PLS_K3_QS=45,PLS_K3_T=128,PLS_K3_MINB=4" 2> gpurun_out/sweep6.err | tee gpurun_out/k2_sweep6_cfg2.jsonl
timeout 900 python tools/k2_sweep.py cfg2 "$V3,PLS_K3_MINB=3,PLS_K2_PHASES=1" 2> gpurun_out/phases_v3_cfg2.txt | tail -1
timeout 900 python tools/k2_sweep.py m512k16 "PLS_K3_QS=0" "PLS_K3_QS=0,PLS_K3_MINB=2" "PLS_K3_QS=0,PLS_K3_T=512" "PLS_K3_QS=0,PLS_K3_T=128,PLS_K3_MINB=4" "PLS_K3_QS=0,PLS_K3_CHAIN=8" 2>> gpurun_out/sweep6.err | tee gpurun_out/k2_sweep6_m512.jsonl
timeout 900 python tools/k2_sweep.py m512k16 "PLS_K3_QS=0,PLS_K2_PHASES=1" 2> gpurun_out/phases_v3_m512.txt | tail -1
tail -5 gpurun_out/sweep6.err
