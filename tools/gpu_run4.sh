#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
PLS_K2_PHASES=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "bench rc=$?"; tail -20 gpurun_out/bench3.err; cut -c1-400 gpurun_out/bench3.json; python -c "
import json; d=json.load(open('gpurun_out/bench3.json')); print(d['stages_ms'], d['value'], d['e2e']['value'])"
