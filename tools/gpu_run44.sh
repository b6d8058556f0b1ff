#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu44.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu44.log
