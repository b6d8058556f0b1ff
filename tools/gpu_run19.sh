#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu19.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu19.log
