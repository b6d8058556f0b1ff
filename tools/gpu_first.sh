#!/bin/bash
# First GPU call: smoke, parity tests, short bench.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python __graft_entry__.py smoke 2>&1 | tail -20
python -m pytest tests -m gpu -x -q 2>&1 | tail -30
