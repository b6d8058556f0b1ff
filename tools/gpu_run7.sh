#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "bnb" 2>&1 | tail -15
timeout 300 python tools/bnb_bench.py 50000 200 16 1e-3 --opt 2>&1 | tail -3 | tee gpurun_out/bnb_a.json
timeout 600 python tools/bnb_bench.py 100000 400 20 0.0 2>&1 | tail -3 | tee gpurun_out/bnb_b.json
