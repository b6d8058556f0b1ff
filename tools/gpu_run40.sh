#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tools/v4_fuzz.py 60 1 > gpurun_out/v4_fuzz.jsonl 2> gpurun_out/v4_fuzz.err; echo "fuzz rc=$?"
tail -2 gpurun_out/v4_fuzz.jsonl | cut -c1-500; tail -3 gpurun_out/v4_fuzz.err
timeout 300 python tools/k2_sweep.py cfg2 '' 2>/dev/null | cut -c1-150
