#!/bin/bash
# A/B of library builds: gpurun_libs/lib_<X>.so copied over the in-tree library one at a time, cfg2 + k20 timings for each
mkdir -p gpurun_out
L=gpurun_out/$1.log
: > $L
cp partitionedls.jl_b200/libpls_cuda.so /tmp/lib_keep.so
for f in gpurun_libs/lib_*.so; do
  echo "== $f" >> $L
  cp $f partitionedls.jl_b200/libpls_cuda.so
  for rep in 1 2; do
    timeout 300 python tools/v5_check.py cfg2 2>&1 | grep -E "cfg2 v5 " | sed -E 's/.*"ms_nnls": ([0-9.]+).*/cfg2 \1/' >> $L
    timeout 300 python tools/v5_check.py k20 2>&1 | grep -E "k20_m200 v5 " | sed -E 's/.*"ms_nnls": ([0-9.]+).*/k20 \1/' >> $L
  done
done
cp /tmp/lib_keep.so partitionedls.jl_b200/libpls_cuda.so
cat $L
