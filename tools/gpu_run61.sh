#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu61.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu61.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench61.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench61.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['stages_ms'])"
timeout 600 python bench.py --steps 3 --warmup 3 --workload cfg2_k17 --no-cpu-baseline > gpurun_out/bench61_k17.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench61_k17.json')); print('k17', d['value'], d['ms_per_step'])"
timeout 600 python bench.py --steps 3 --warmup 3 --workload k20_m200 --no-cpu-baseline > gpurun_out/bench61_k20.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench61_k20.json')); print('k20', d['value'], d['ms_per_step'])"
