#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "predict or toy or multi" > gpurun_out/pytest_gpu38.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu38.log
python - <<'PY'
import time, json, numpy as np, sys, os
sys.path.insert(0, os.getcwd())
import __graft_entry__ as g
pkg = g.load_package()
from importlib import import_module
synth = import_module(g.PKG_NAME + ".synth")
X, y, P, eta = synth.make_config("cfg2")
ctx = pkg.Context(0); ctx.load(X, y, P, eta=eta)
w = np.random.default_rng(0).standard_normal(X.shape[1] + 1)
for _ in range(3):
    t0 = time.perf_counter(); yh = ctx.predict_resident(w, len(y)); dt = time.perf_counter() - t0
ms = ctx.stats()["ms_recompute"]
ref = X @ w[:-1] + w[-1]
print(json.dumps(dict(kernel="k7_predict_rows", N=len(y), M=X.shape[1], ms_kernel=ms, gbs=(8.0 * X.size + 8 * len(y)) / (ms * 1e-3) / 1e9, s_call=dt,
                      max_rel_err=float(np.abs(yh - ref).max() / np.abs(ref).max()))))
PY
