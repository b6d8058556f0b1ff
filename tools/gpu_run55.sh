#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stalled or paired or predict" > gpurun_out/pytest_gpu55.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu55.log
timeout 1500 python tools/v4_fuzz.py 150 777 > gpurun_out/v4_fuzz4.jsonl 2> gpurun_out/v4_fuzz4.err; echo "fuzz rc=$?"
tail -1 gpurun_out/v4_fuzz4.jsonl | cut -c1-300; tail -2 gpurun_out/v4_fuzz4.err
