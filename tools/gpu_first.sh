#!/bin/bash
# First GPU call: FP64 peaks, smoke, parity tests.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
./tools/fp64_peak.bin | tee gpurun_out/fp64_peak.json
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -20
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -40
