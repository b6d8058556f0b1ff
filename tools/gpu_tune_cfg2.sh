#!/bin/bash
# window size / walks per SM on cfg2 (K = 16, groups of 12-13)
mkdir -p gpurun_out
L=gpurun_out/$1.log
: > $L
run() { echo "== $*" >> $L; env "$@" timeout 300 python tools/v5_check.py cfg2 2>&1 | grep -E "cfg2 v5 " | sed -E 's/.*"ms_nnls": ([0-9.]+).*"sweeps": ([0-9]+).*"occ": ([0-9]+).*/cfg2 \1 sweeps \2 occ \3/' >> $L; }
run X=1
run PLS_K5_NR=80
run PLS_K5_NR=80 PLS_K5_MARGIN=8
run PLS_K5_NR=72
run PLS_K5_NR=96 PLS_K5_MARGIN=16
cat $L
