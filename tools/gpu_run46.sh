#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tools/v4_fuzz.py 80 7 > gpurun_out/v4_fuzz2.jsonl 2> gpurun_out/v4_fuzz2.err; echo "fuzz rc=$?"
tail -1 gpurun_out/v4_fuzz2.jsonl | cut -c1-300; tail -3 gpurun_out/v4_fuzz2.err
