"""K1 (Gram build) alone at the BASELINE shapes: wall time of pls_gram_build (+ stream sync) on a
resident data set, algorithmic FLOPs N*M'(M'+1) + 2NM' + 2N against the measured DMMA peak."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
peak = json.load(open(os.path.join(g.ROOT, "profiles", "fp64_peaks_r01.json")))["dmma_m8n8k4_tflops"]
ctx = pkg.Context(0)
for name, N, M, K in [("cfg2", 100_000, 200, 16), ("cfg3", 1_000_000, 512, 24), ("cfg5", 500_000, 800, 32)]:
    rng = np.random.default_rng(1)
    X = np.asfortranarray(rng.standard_normal((M, N), dtype=np.float64).T)   # column-major N x M
    y = rng.standard_normal(N)
    P = np.zeros((M, K), dtype=np.int64); P[np.arange(M), (np.arange(M) * K) // M] = 1
    ctx.load(X, y, P, eta=1e-3)
    ts = []
    for rep in range(6):
        t0 = time.perf_counter(); ctx.gram_build(); ctx.gram_raw(); ts.append(time.perf_counter() - t0)
    t = min(ts[1:])
    Mp = M + 1
    flops = N * Mp * (Mp + 1) + 2 * N * Mp + 2 * N
    rec = dict(shape=name, N=N, M=M, ms=t * 1e3, tflops=flops / t / 1e12, frac_of_dmma_peak=flops / t / 1e12 / peak,
               hbm_floor_ms=8.0 * N * (M + 2) / 6553e9 * 1e3)
    print(json.dumps(rec), flush=True)
    del X
