#!/bin/bash
# default dispatch on wide problems: m512_k16, m512_k20 and one rank's 1/8 share of a K = 24 enumeration (v5 against v4)
mkdir -p gpurun_out
L=gpurun_out/$1.log
timeout 1200 python - > $L 2>&1 <<'PY'
import sys, os, json
sys.path.insert(0, '.')
import __graft_entry__ as g
pkg = g.load_package()
from importlib import import_module
synth = import_module(g.PKG_NAME + ".synth")
ctx = pkg.Context(0)
for name, cnt in (("m512_k16", None), ("m512_k20", None), ("m512_k24", 1 << 21)):
    N, M, K, eta, seed, mixed = synth.CONFIGS[name]
    X, y, P = synth.make_synthetic(N, M, K, seed, mixed_sign=mixed)
    ctx.load(X, y, P, eta=eta); ctx.gram_build(); ctx.gram_finalize()
    res = {}
    for label, e in (("default", {}), ("v4", dict(PLS_K2_NO_V5="1"))):
        for k in ("PLS_K2_IMPL", "PLS_K2_NO_V5"): os.environ.pop(k, None)
        os.environ.update(e)
        for _ in range(2):
            r = ctx.opt_solve_pairs(0, cnt if cnt else (1 << K))
        st = ctx.stats()
        res[label] = (r["b_best"], r["obj_gram"])
        print(name, label, json.dumps(dict(variant=st["k2_variant"], T=st["k2_threads"], grid=st["k2_grid"], ms=st["ms_nnls"], b=r["b_best"], sweeps=st["pivots"], iters=st["bpp_iters"], rebuilds=st["rebuilds"], blocked=st["blocked"], spills=st["spills"], drift=st["k2_max_drift"])), flush=True)
    print(name, "same winner", res["default"][0] == res["v4"][0] and abs(res["default"][1] - res["v4"][1]) <= 1e-9 * res["v4"][1], flush=True)
PY
cat $L
