#!/bin/bash
# small tuning sweep of the v5 launch parameters on k20_m200 (and cfg2)
mkdir -p gpurun_out
L=gpurun_out/$1.log
: > $L
run() { echo "== $*" >> $L; env "$@" timeout 300 python tools/v5_check.py k20 2>&1 | grep -E "k20_m200 v5 " | sed -E 's/.*"ms_nnls": ([0-9.]+).*"sweeps": ([0-9]+).*/k20 \1 sweeps \2/' >> $L; }
run X=1
run PLS_K5_MARGIN=10
run PLS_K5_MARGIN=14
run PLS_K5_MARGIN=16
run PLS_K5_L=5
run PLS_K5_L=4
run PLS_K5_VERIFY=256
run PLS_K5_VERIFY=64
run PLS_K5_NR=96
cat $L
