#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tools/v4_fuzz.py 200 12345 > gpurun_out/v4_fuzz3.jsonl 2> gpurun_out/v4_fuzz3.err; echo "fuzz rc=$?"
tail -1 gpurun_out/v4_fuzz3.jsonl | cut -c1-300; tail -3 gpurun_out/v4_fuzz3.err
