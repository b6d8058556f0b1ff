#!/bin/bash
# K2 v5 at 1..5 walks per SM on k20_m200: is a walk's latency or the SM's throughput the limit?
mkdir -p gpurun_out
L=gpurun_out/$1.log
: > $L
for o in 1 2 3 4 5; do
  echo "== occ $o" >> $L
  PLS_K5_OCC=$o timeout 300 python tools/v5_check.py k20 2>&1 | grep -E "k20_m200 v5 " >> $L
done
cat $L
