#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_v3.err; cut -c1-600 gpurun_out/bench_v3.json
timeout 600 python bench.py --steps 3 --warmup 3 --workload k20_m200 --no-cpu-baseline > gpurun_out/bench_k20.json 2> gpurun_out/bench_k20.err; echo "k20 rc=$?"; cut -c1-300 gpurun_out/bench_k20.json
timeout 600 python tools/alt_bench.py 200000 1000 50 4096 2>&1 | tail -2 | tee gpurun_out/alt_cfg4.json
timeout 600 python tools/bnb_bench.py 200000 800 32 0.0 2>&1 | tail -2 | tee gpurun_out/bnb_cfg5.json
