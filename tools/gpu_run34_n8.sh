#!/bin/bash
# BASELINE configs[3] (Alt, N=1M, M=1000, K=50, 4096 restarts) and configs[4] (BnB, N=500k, M=800, K=32) at full size on 8 GPUs
mkdir -p gpurun_out
NGPU=8 timeout 600 python tools/alt_bench.py 1000000 1000 50 4096 2>&1 | tail -1 | tee gpurun_out/cfg4_alt_8gpu.json
NGPU=8 BNB_REPS=2 BNB_FLIP=0.2 PLS_BNB_MAX_NODES=2000000 timeout 400 python tools/bnb_bench.py 500000 800 32 0.0 2>&1 | tail -1 | tee gpurun_out/cfg5_bnb_8gpu.json
