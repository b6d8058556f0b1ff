#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q -k "multi_gpu or toy" 2>&1 | tail -15
python - <<'PY'
import sys, time, json
sys.path.insert(0, '.')
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
from importlib import import_module
synth = import_module(g.PKG_NAME + ".synth")
X, y, P, eta = synth.make_config("cfg2_k17")
for devs in ([0], [0, 1]):
    c = pkg.Context(devs)
    c.load(X, y, P, eta=eta)
    ts = []
    for rep in range(4):
        t0 = time.perf_counter(); r = c.opt_fit_resident(); ts.append(time.perf_counter() - t0)
    t0 = time.perf_counter(); r2 = c.opt_fit(X, y, P, eta=eta); te = time.perf_counter() - t0
    print(json.dumps(dict(devs=devs, ms_resident=min(ts) * 1e3, ms_from_host=te * 1e3, b=r["b_best"], opt=r["opt"], solves_per_s=(1 << 18) / min(ts),
                          ms_gram=r["stats"]["ms_gram"], ms_nnls=r["stats"]["ms_nnls"])), flush=True)
    c.close()
PY
