#!/bin/bash
# ncu --set full capture of the wide v5 variant (512-thread walks) on a 2^18-problem range of m512_k20
mkdir -p gpurun_out
cat > /tmp/wide_run.py <<'PY'
import sys, os
sys.path.insert(0, '.')
import __graft_entry__ as g
pkg = g.load_package()
from importlib import import_module
synth = import_module(g.PKG_NAME + ".synth")
N, M, K, eta, seed, mixed = synth.CONFIGS["m512_k20"]
X, y, P = synth.make_synthetic(N, M, K, seed, mixed_sign=mixed)
ctx = pkg.Context(0)
ctx.load(X, y, P, eta=eta); ctx.gram_build(); ctx.gram_finalize()
for _ in range(2):
    r = ctx.opt_solve_pairs(0, 1 << 18)
st = ctx.stats()
print({k: st[k] for k in ("k2_variant", "k2_threads", "ms_nnls", "pivots", "grad_evals", "sum_p", "bpp_iters")})
PY
python /tmp/wide_run.py > gpurun_out/$1_plain.log 2>&1 || { tail -3 gpurun_out/$1_plain.log; exit 1; }
tail -1 gpurun_out/$1_plain.log
ncu --set full --clock-control none --import-source on -k regex:k2v5 -s 1 -c 1 -f -o gpurun_out/$1 python /tmp/wide_run.py > gpurun_out/$1_ncu.log 2>&1
ls -la gpurun_out/$1.ncu-rep
