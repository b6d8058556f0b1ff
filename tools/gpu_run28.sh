#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu28.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu28.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench28.json 2> gpurun_out/bench28.err; echo "bench rc=$?"
cut -c1-1200 gpurun_out/bench28.json
timeout 600 python bench.py --steps 3 --warmup 3 --workload k20_m200 --no-cpu-baseline > gpurun_out/bench28_k20.json 2>> gpurun_out/bench28.err
cut -c1-400 gpurun_out/bench28_k20.json
timeout 600 python bench.py --steps 3 --warmup 3 --workload m512_k16 --no-cpu-baseline > gpurun_out/bench28_m512.json 2>> gpurun_out/bench28.err
cut -c1-400 gpurun_out/bench28_m512.json
