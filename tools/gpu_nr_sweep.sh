#!/bin/bash
# K2 v5 window sizes / walks per SM on k20_m200
mkdir -p gpurun_out
L=gpurun_out/$1.log
: > $L
for cfg in "64 12" "64 8" "72 12" "72 10" "72 8" "80 12"; do
  set -- $cfg
  echo "== NR $1 margin $2" >> $L
  PLS_K5_NR=$1 PLS_K5_MARGIN=$2 timeout 300 python tools/v5_check.py k20 2>&1 | grep -E "k20_m200 v5 |same" >> $L
done
cat $L
