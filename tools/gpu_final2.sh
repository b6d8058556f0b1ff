#!/bin/bash
# end-of-round measurements (two-level K2): tests, smoke, bench lines, ncu launch list, full capture of the dominant kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/final2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final2_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final2_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/final2_bench.json 2> gpurun_out/final2_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final2_bench_reference.json 2>> gpurun_out/final2_bench.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 3 --warmup 3 --workload k20_m200 --no-cpu-baseline > gpurun_out/final2_bench_k20.json 2>/dev/null; echo "k20 rc=$?"
timeout 900 python bench.py --steps 3 --warmup 3 --workload m512_k16 --no-cpu-baseline > gpurun_out/final2_bench_m512k16.json 2>/dev/null; echo "m512 rc=$?"
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/final2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final2_launches.csv $CMD > gpurun_out/final2_ncu_list.log 2>&1
$CMD > gpurun_out/final2_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k2v4_orthant -s 1 -c 1 -f -o gpurun_out/final2_k2_prof $CMD > gpurun_out/final2_ncu_full.log 2>&1
tail -2 gpurun_out/final2_ncu_full.log
python -c "
import json
for f in ['final2_bench','final2_bench_reference','final2_bench_k20','final2_bench_m512k16']:
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d.get('ms_per_step'), d.get('e2e',{}).get('value'), d.get('roofline',{}).get('frac'), d.get('cpu_baseline'))"
