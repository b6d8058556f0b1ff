#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu43_n2.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu43_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench43_n2.json 2> gpurun_out/bench43_n2.err; echo "bench n2 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench43_n2.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['result'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench43_ref_n2.json 2>> gpurun_out/bench43_n2.err; echo "ref n2 rc=$?"; cut -c1-200 gpurun_out/bench43_ref_n2.json
