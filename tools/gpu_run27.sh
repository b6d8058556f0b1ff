#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "v4 or toy" > gpurun_out/pytest_gpu27.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu27.log
timeout 900 python tools/k2_sweep.py cfg2 '' 'PLS_K4_L=6' 'PLS_K4_L=4' 'PLS_K2_PHASES=1' > gpurun_out/k2_sweep27_cfg2.jsonl 2> gpurun_out/sweep27.err; echo "sweep rc=$?"
cut -c1-200 gpurun_out/k2_sweep27_cfg2.jsonl
tail -24 gpurun_out/sweep27.err
timeout 600 python tools/k2_sweep.py m512k16 'PLS_K4_L=5' 'PLS_K4_L=6' 'PLS_K4_L=7' 'PLS_K4_L=6,PLS_K2_PHASES=1' > gpurun_out/k2_sweep27_m512.jsonl 2> gpurun_out/sweep27m.err
cut -c1-200 gpurun_out/k2_sweep27_m512.jsonl
tail -24 gpurun_out/sweep27m.err
