#!/bin/bash
mkdir -p gpurun_out
python tools/debug_case.py tests/golden/stall_n144_m117_rho09.npz 'PLS_K2_IMPL=v3' 'PLS_K2_IMPL=v4,PLS_K4_GRID=4' '' 2>&1 | tail -4
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu54.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu54.log
timeout 1500 python tools/v4_fuzz.py 200 12345 > gpurun_out/v4_fuzz3.jsonl 2> gpurun_out/v4_fuzz3.err; echo "fuzz rc=$?"
tail -1 gpurun_out/v4_fuzz3.jsonl | cut -c1-300; tail -2 gpurun_out/v4_fuzz3.err
