#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu51.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu51.log
timeout 1500 python tools/v4_fuzz.py 200 12345 > gpurun_out/v4_fuzz3.jsonl 2> gpurun_out/v4_fuzz3.err; echo "fuzz rc=$?"
tail -1 gpurun_out/v4_fuzz3.jsonl | cut -c1-300; tail -2 gpurun_out/v4_fuzz3.err
timeout 300 python tools/k2_sweep.py cfg2 '' 'PLS_K2_IMPL=v3' 2>/dev/null | cut -c1-130
