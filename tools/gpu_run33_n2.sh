#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "multi_gpu or bnb" > gpurun_out/pytest_gpu33.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu33.log
BNB_REPS=2 BNB_FLIP=0.2 PLS_BNB_MAX_NODES=1000000 timeout 300 python tools/bnb_bench.py 200000 800 32 0.0 2>&1 | tail -1 | tee gpurun_out/bnb33_n1.json
NGPU=2 BNB_REPS=2 BNB_FLIP=0.2 PLS_BNB_MAX_NODES=1000000 timeout 300 python tools/bnb_bench.py 200000 800 32 0.0 2>&1 | tail -1 | tee gpurun_out/bnb33_n2.json
BNB_REPS=2 PLS_BNB_MAX_NODES=2000000 timeout 300 python tools/bnb_bench.py 100000 400 20 0.0 2>&1 | tail -1 | tee gpurun_out/bnb33_mixed_n1.json
NGPU=2 BNB_REPS=2 PLS_BNB_MAX_NODES=2000000 timeout 300 python tools/bnb_bench.py 100000 400 20 0.0 2>&1 | tail -1 | tee gpurun_out/bnb33_mixed_n2.json
