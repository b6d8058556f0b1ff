#!/bin/bash
# (round 2: compute-sanitizer is closed on this pool -- the plain run of the small cases below still checks them against the oracle)
mkdir -p gpurun_out
python tools/sanitize_small.py > gpurun_out/sanitize_plain.log 2>&1 || { tail -5 gpurun_out/sanitize_plain.log; exit 1; }
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py ${SAN_MODES:-opt opth optv5 bnb alt} > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok|Error|hazard" gpurun_out/sanitize_$tool.log | head -12
done
