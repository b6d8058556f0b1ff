#!/bin/bash
# end-of-round measurements: bench lines, ncu launch list and full capture of the dominant kernel
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3"
timeout 900 $B > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --steps 3 --warmup 3 --workload k20_m200 --no-cpu-baseline > gpurun_out/final_bench_k20.json 2>/dev/null; echo "k20 rc=$?"
timeout 900 python bench.py --steps 3 --warmup 3 --workload m512_k16 --no-cpu-baseline > gpurun_out/final_bench_m512k16.json 2>/dev/null; echo "m512 rc=$?"
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k2v3_orthant -s 1 -c 1 -o gpurun_out/final_k2_prof $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
BNB_REPS=2 BNB_FLIP=0.2 PLS_BNB_MAX_NODES=1000000 timeout 300 python tools/bnb_bench.py 500000 800 32 0.0 2>&1 | tail -1 | tee gpurun_out/final_bnb_cfg5.json
timeout 300 python tools/alt_bench.py 200000 1000 50 4096 2>&1 | tail -1 | tee gpurun_out/final_alt_cfg4.json
cut -c1-300 gpurun_out/final_bench.json
