#!/bin/bash
# cold-start variants of the v5 K2 kernel: correctness on the small cases, then cold-only launches and the two full configurations
mkdir -p gpurun_out
L=gpurun_out/$1.log
: > $L
timeout 600 python tools/v5_check.py small 2>&1 | grep -E '"ok": false|RESULT' >> $L
for f in 1 0; do
  echo "== fused $f" >> $L
  PLS_K5_FUSED=$f timeout 300 python tools/v5_check.py cold 2>&1 | grep -E "^cold|RESULT" >> $L
done
timeout 300 python tools/v5_check.py cfg2 2>&1 | grep -E "v5|same" >> $L
timeout 300 python tools/v5_check.py k20 2>&1 | grep -E "v5|same" >> $L
tail -40 $L
