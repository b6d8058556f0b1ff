#!/bin/bash
# ncu --set full capture of the K2 v5 kernel on a range so short that only the cold start runs (1024 patterns of k20_m200)
mkdir -p gpurun_out
export PLS_K5_MARGIN=${PLS_K5_MARGIN:-10}
ncu --set full --clock-control none --import-source on -k regex:k2v5 -s 3 -c 1 -f -o gpurun_out/$1 python tools/v5_check.py cold > gpurun_out/$1.log 2>&1
tail -2 gpurun_out/$1.log
