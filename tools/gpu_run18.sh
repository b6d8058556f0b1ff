#!/bin/bash
mkdir -p gpurun_out
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/final_bench_n$n.json 2> gpurun_out/final_bench_n$n.err; echo "n$n rc=$?"; tail -1 gpurun_out/final_bench_n$n.err | cut -c1-200; cut -c1-200 gpurun_out/final_bench_n$n.json
done
