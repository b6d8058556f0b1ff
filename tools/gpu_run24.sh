#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/k2_sweep.py cfg2 'PLS_K2_IMPL=v3' '' 'PLS_K4_L=6' 'PLS_K4_L=7' 'PLS_K4_L=8' 'PLS_K4_L=6,PLS_K4_QS=21' 'PLS_K4_L=6,PLS_K4_QS=36' 'PLS_K4_L=6,PLS_K4_MINB=5' 'PLS_K4_L=6,PLS_K4_MINB=3' 'PLS_K4_L=6,PLS_K4_T=256,PLS_K4_MINB=2' 'PLS_K4_L=6,PLS_K2_PHASES=1' > gpurun_out/k2_sweep24_cfg2.jsonl 2> gpurun_out/sweep24.err; echo "sweep rc=$?"
cut -c1-200 gpurun_out/k2_sweep24_cfg2.jsonl
tail -24 gpurun_out/sweep24.err
timeout 600 python tools/k2_sweep.py m512k16 'PLS_K2_IMPL=v3' '' 'PLS_K4_L=3' 'PLS_K4_L=4' > gpurun_out/k2_sweep24_m512.jsonl 2>> gpurun_out/sweep24.err
cut -c1-200 gpurun_out/k2_sweep24_m512.jsonl
ncu --set full --clock-control none --import-source on -k regex:k2v4_orthant -c 1 -f -o gpurun_out/k2_prof24 python tools/k2_sweep.py cfg2 'PLS_K4_L=6' > gpurun_out/ncu24.log 2>&1
tail -3 gpurun_out/ncu24.log
