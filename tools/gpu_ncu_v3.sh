#!/bin/bash
# full ncu capture of the v3 K2 kernel at cfg2 (after the plain command exits 0)
mkdir -p gpurun_out
export PLS_K2_IMPL=v3 PLS_K3_QS=0 PLS_K3_MINB=3
CMD="python tools/k2_sweep.py cfg2 PLS_K2_IMPL=v3,PLS_K3_QS=0,PLS_K3_MINB=3"
$CMD > gpurun_out/plain_v3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k2v3_orthant -s 1 -c 1 -o gpurun_out/k2v3_prof $CMD > gpurun_out/ncu_v3.log 2>&1
tail -3 gpurun_out/ncu_v3.log
