#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tools/cfg3_multi.py 1000000 512 24 8 > gpurun_out/cfg3_8gpu_pairs.json 2> gpurun_out/cfg3_8gpu_pairs.err; echo "cfg3 rc=$?"
cat gpurun_out/cfg3_8gpu_pairs.json
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/bench45_n$n.json 2> gpurun_out/bench45_n$n.err; echo "n=$n rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench45_n$n.json')); print($n, d['value'], d['ms_per_step'], d['e2e']['value'], d['stages_ms'], d['result'])"
done
