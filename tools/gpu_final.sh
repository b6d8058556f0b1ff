#!/bin/bash
# end-of-round evidence on one GPU: test suite, three fuzz seeds, the default bench line, the reference arm, ncu captures
mkdir -p gpurun_out
T=$1
python -m pytest tests -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -2 gpurun_out/${T}_tests.log
for s in 11 12 13; do python tools/k2_fuzz.py 150 $s > gpurun_out/${T}_fuzz_s$s.log 2>&1; tail -1 gpurun_out/${T}_fuzz_s$s.log; done
python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 600 gpurun_out/${T}_bench.json | head -c 300; echo
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; head -c 200 gpurun_out/${T}_bench_ref.json; echo
tools/gpu_ncu_k2.sh k2v5 ${T}_ncu
