#!/bin/bash
# first run of the two-level K2 path: parity tests for the v4 variants, then a timing sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "v4 or toy or full_size" > gpurun_out/pytest_gpu23.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu23.log
timeout 600 python tools/k2_sweep.py cfg2 'PLS_K2_IMPL=v3' '' 'PLS_K4_L=4' 'PLS_K4_L=6' 'PLS_K4_L=4,PLS_K2_PHASES=1' > gpurun_out/k2_sweep23_cfg2.jsonl 2> gpurun_out/sweep23.err; echo "sweep rc=$?"
cut -c1-400 gpurun_out/k2_sweep23_cfg2.jsonl
tail -30 gpurun_out/sweep23.err
