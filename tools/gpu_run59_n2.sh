#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/final4_bench.json 2> gpurun_out/final4_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final4_bench_reference.json 2>> gpurun_out/final4_bench.err; echo "ref rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/final4_bench_n2.json 2>> gpurun_out/final4_bench.err; echo "n2 rc=$?"
python -c "
import json
for f in ['final4_bench','final4_bench_reference','final4_bench_n2']:
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['unit'], d.get('ms_per_step'), d.get('e2e',{}).get('value'))"
