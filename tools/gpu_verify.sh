#!/bin/bash
# full verification of a kernel change: GPU test suite, one fuzz seed, default bench line
#   tools/gpu_verify.sh <tag> [fuzz-seed]
mkdir -p gpurun_out
T=$1; S=${2:-21}
python -m pytest tests -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -3 gpurun_out/${T}_tests.log
python tools/k2_fuzz.py 100 $S > gpurun_out/${T}_fuzz_s$S.log 2>&1; tail -1 gpurun_out/${T}_fuzz_s$S.log
python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
j=json.loads([l for l in open("gpurun_out/${T}_bench.json") if l.startswith("{")][-1])
print(j["value"], j["ms_per_step"], j["e2e"]["ms_per_step"], j["stages_ms"], j["roofline"]["frac"])
print({k:(v if not isinstance(v,dict) else {a:b for a,b in v.items() if "ms" in a}) for k,v in j["extra"]["cfg2"].items()})
PY
