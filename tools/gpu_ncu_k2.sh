#!/bin/bash
# ncu launch list + full capture of the dominant K2 kernel (run only after the plain command exits 0)
#   tools/gpu_ncu_k2.sh <kernel-regex> <out-name> <bench args...>
mkdir -p gpurun_out
KRE=$1; OUT=$2; shift 2
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras $@"
$CMD > gpurun_out/${OUT}_plain.json 2> gpurun_out/${OUT}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/${OUT}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${OUT}_launches.csv $CMD > gpurun_out/${OUT}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:$KRE -s 1 -c 1 -f -o gpurun_out/${OUT} $CMD > gpurun_out/${OUT}_ncu_full.log 2>&1
tail -3 gpurun_out/${OUT}_ncu_full.log
ls -la gpurun_out/${OUT}.ncu-rep
