#!/bin/bash
# full ncu capture of the default K2 kernel at cfg2 (after the plain command exits 0)
mkdir -p gpurun_out
S="${PLS_SWEEP_SETTING:-PLS_K3_T=128}"
python tools/k2_sweep.py cfg2 "$S" > gpurun_out/plain22.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k2v3_orthant -s 1 -c 1 -f -o gpurun_out/k2_prof22 python tools/k2_sweep.py cfg2 "$S" > gpurun_out/ncu22.log 2>&1
tail -3 gpurun_out/ncu22.log
