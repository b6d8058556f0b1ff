#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench29.json 2> gpurun_out/bench29.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench29.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['clocks'], d['roofline']['frac'], d['roofline']['traffic'], d['stages_ms'])"
SWEEP_COUNT_LOG2=20 timeout 900 python tools/k2_sweep.py m512k24 'PLS_K2_IMPL=v3' 'PLS_K4_L=5' 'PLS_K4_L=6' 'PLS_K4_L=7' 'PLS_K4_L=8' 'PLS_K4_L=6,PLS_K4_MINB=3' 'PLS_K4_L=7,PLS_K4_T=128,PLS_K4_MINB=3' 'PLS_K4_L=7,PLS_K2_PHASES=1' > gpurun_out/k2_sweep29_m512k24.jsonl 2> gpurun_out/sweep29.err; echo "sweep rc=$?"
cut -c1-220 gpurun_out/k2_sweep29_m512k24.jsonl
tail -24 gpurun_out/sweep29.err
SWEEP_COUNT_LOG2=21 timeout 900 python tools/k2_sweep.py k20 'PLS_K2_IMPL=v3' '' 'PLS_K4_L=6' 'PLS_K4_L=7' 'PLS_K4_L=6,PLS_K2_PHASES=1' > gpurun_out/k2_sweep29_k20.jsonl 2> gpurun_out/sweep29b.err
cut -c1-220 gpurun_out/k2_sweep29_k20.jsonl
tail -24 gpurun_out/sweep29b.err
