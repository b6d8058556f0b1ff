#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1500 python tools/cfg3_multi.py 1000000 512 24 2>&1 | tail -2 | tee gpurun_out/cfg3_8gpu.json
