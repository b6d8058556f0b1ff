#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"; tail -5 gpurun_out/bench_n2.err; cut -c1-600 gpurun_out/bench_n2.json
python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print(d['stages_ms'], d['value'], d['e2e'], d['result'], d['gpu_launches'])"
python bench.py --gpus 1 --steps 3 --warmup 3 --workload cfg2_k17 --no-cpu-baseline > gpurun_out/bench_k17_n1.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_k17_n1.json')); print('1gpu k17', d['stages_ms'], d['value'], d['e2e']['value'], d['result'])"
