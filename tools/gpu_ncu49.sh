#!/bin/bash
# ncu --set full captures of the BnB wave kernel (K5), the Alt restart kernel (K6) and the row-streaming kernel (K4/K7)
mkdir -p gpurun_out
export BNB_REPS=1 BNB_FLIP=0.2 PLS_BNB_MAX_NODES=1000000
python tools/bnb_bench.py 200000 800 32 0.0 > gpurun_out/plain49_bnb.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k5_bnb_expand -s 20 -c 1 -f -o gpurun_out/k5_prof49 python tools/bnb_bench.py 200000 800 32 0.0 > gpurun_out/ncu49_bnb.log 2>&1
tail -1 gpurun_out/ncu49_bnb.log
python tools/alt_bench.py 100000 1000 50 1024 > gpurun_out/plain49_alt.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k6_alt -c 1 -f -o gpurun_out/k6_prof49 python tools/alt_bench.py 100000 1000 50 1024 > gpurun_out/ncu49_alt.log 2>&1
tail -1 gpurun_out/ncu49_alt.log
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain49_k4.log 2>&1 &&
ncu --set full --clock-control none -k regex:k47_rows -c 1 -f -o gpurun_out/k4_prof49 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu49_k4.log 2>&1
tail -1 gpurun_out/ncu49_k4.log
