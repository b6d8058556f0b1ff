#!/bin/bash
# strong-scaling bench lines on one multi-GPU box: N = 2, 4, 8 (torchrun, one rank per GPU), plus the multi-GPU tests
mkdir -p gpurun_out
T=$1
python -m pytest tests -x -q -m gpu -k "multi or dist or shard or devices" > gpurun_out/${T}_mgpu_tests.log 2>&1; tail -2 gpurun_out/${T}_mgpu_tests.log
for n in ${NS:-2 4 8}; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/${T}_n$n.json 2> gpurun_out/${T}_n$n.err
  python - <<PY
import json
try:
    j=json.loads([l for l in open("gpurun_out/${T}_n$n.json") if l.startswith("{")][-1])
    ex=j.get("extra",{})
    print($n, j["value"], j["ms_per_step"], j["e2e"]["ms_per_step"], j["stages_ms"], {k:(v.get("ms_per_step_resident") if isinstance(v,dict) else v) for k,v in ex.items()})
except Exception as e: print("N=$n failed", e)
PY
done
