#!/bin/bash
mkdir -p gpurun_out
python tools/k2_sweep.py cfg2 '' > gpurun_out/plain26.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k2v4_orthant -c 1 -f -o gpurun_out/k2_prof26 python tools/k2_sweep.py cfg2 '' > gpurun_out/ncu26.log 2>&1
tail -2 gpurun_out/ncu26.log
