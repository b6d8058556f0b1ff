#!/bin/bash
# v5 with large windows (256-thread walks) against v4 on wide problems
mkdir -p gpurun_out
L=gpurun_out/$1.log
: > $L
for cfg in ${CFGS:-m512_k16}; do
for nr in ${NRS:-128 160 192}; do
  echo "== $cfg v5 NR $nr" >> $L
  PLS_K5_NR=$nr timeout 600 python - $cfg >> $L 2>&1 <<'PY'
import sys, os, json
sys.path.insert(0, '.')
import __graft_entry__ as g
pkg = g.load_package()
from importlib import import_module
synth = import_module(g.PKG_NAME + ".synth")
name = sys.argv[1]
N, M, K, eta, seed, mixed = synth.CONFIGS[name]
X, y, P = synth.make_synthetic(N, M, K, seed, mixed_sign=mixed)
ctx = pkg.Context(0)
ctx.load(X, y, P, eta=eta)
out = {}
for label, e in (("v5", dict(PLS_K2_IMPL="v5")), ("v4", dict(PLS_K2_NO_V5="1"))):
    for k in ("PLS_K2_IMPL", "PLS_K2_NO_V5"): os.environ.pop(k, None)
    os.environ.update(e)
    if label == "v4" and os.environ.get("PLS_K5_NR") != os.environ.get("_FIRST_NR", os.environ.get("PLS_K5_NR")): pass
    for _ in range(2): r = ctx.opt_fit_resident()
    st = r["stats"]
    out[label] = (r["b_best"], r["opt"])
    print(label, json.dumps(dict(variant=st["k2_variant"], T=st["k2_threads"], occ=st["k2_ctas_per_sm"], grid=st["k2_grid"], ms=st["ms_nnls"], b=r["b_best"], sweeps=st["pivots"], streams=st["grad_evals"], sum_s=st["sum_p"], iters=st["bpp_iters"], rebuilds=st["rebuilds"], blocked=st["blocked"], spills=st["spills"], drift=st["k2_max_drift"])), flush=True)
print("same winner", out["v5"][0] == out["v4"][0] and abs(out["v5"][1] - out["v4"][1]) <= 1e-9 * out["v4"][1])
PY
done; done
cat $L
