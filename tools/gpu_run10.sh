#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/alt_bench.py 200000 1000 50 4096 2>&1 | tail -2 | tee gpurun_out/alt_cfg4.json
BNB_REPS=1 BNB_FLIP=0.2 PLS_BNB_MAX_NODES=300000 timeout 300 python tools/bnb_bench.py 200000 800 32 0.0 2>&1 | tail -2 | tee gpurun_out/bnb_cfg5_flip02.json
BNB_REPS=1 PLS_BNB_MAX_NODES=300000 timeout 300 python tools/bnb_bench.py 100000 480 24 0.0 2>&1 | tail -2 | tee gpurun_out/bnb_k24_mixed.json
