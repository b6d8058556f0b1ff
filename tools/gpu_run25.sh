#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "v4 or v3g or toy" > gpurun_out/pytest_gpu25.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu25.log
timeout 900 python tools/k2_sweep.py cfg2 'PLS_K2_IMPL=v3' '' 'PLS_K4_L=6' 'PLS_K4_L=5,PLS_K4_QS=15' 'PLS_K4_L=5,PLS_K4_QS=21' 'PLS_K4_L=6,PLS_K4_QS=21' 'PLS_K4_L=6,PLS_K4_QS=28' 'PLS_K4_L=7,PLS_K4_QS=28' 'PLS_K4_L=6,PLS_K4_MINB=5' > gpurun_out/k2_sweep25_cfg2.jsonl 2> gpurun_out/sweep25.err; echo "sweep rc=$?"
cut -c1-200 gpurun_out/k2_sweep25_cfg2.jsonl
timeout 600 python tools/k2_sweep.py m512k16 'PLS_K2_IMPL=v3' 'PLS_K4_L=4' 'PLS_K4_L=5' 'PLS_K4_L=6' 'PLS_K4_L=5,PLS_K4_T=128,PLS_K4_MINB=3' > gpurun_out/k2_sweep25_m512.jsonl 2>> gpurun_out/sweep25.err
cut -c1-200 gpurun_out/k2_sweep25_m512.jsonl
