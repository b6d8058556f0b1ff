#!/bin/bash
# 2-GPU validation: all GPU tests (incl. the in-library multi-GPU test), smoke, 2-rank bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/final_bench_n2.json 2> gpurun_out/final_bench_n2.err; echo "n2 rc=$?"; tail -2 gpurun_out/final_bench_n2.err; cut -c1-250 gpurun_out/final_bench_n2.json
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/final_bench_n1b.json 2>/dev/null; cut -c1-250 gpurun_out/final_bench_n1b.json
