#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tools/v4_fuzz.py 120 99 > gpurun_out/v4_fuzz5.jsonl 2> gpurun_out/v4_fuzz5.err; echo "fuzz rc=$?"; tail -1 gpurun_out/v4_fuzz5.jsonl | cut -c1-200; tail -2 gpurun_out/v4_fuzz5.err
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu58.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu58.log
