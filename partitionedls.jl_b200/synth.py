"""Synthetic workloads of the shapes BASELINE.json names: X ~ N(0,1) i.i.d. FP64
column-major, contiguous near-equal groups, w_m = |z_m| * s_group (mixed-sign for the BnB config),
y = X w + 0.5 + N(0,1) noise.  Seeds are 20240414 + config index."""
from __future__ import annotations

import numpy as np

CONFIGS = {
    # name: (N, M, K, eta, seed, mixed_sign)
    "cfg1_toy": None,
    "cfg2": (100_000, 200, 16, 1e-3, 20240416, False),
    "cfg2_k17": (100_000, 200, 17, 1e-3, 20240416, False),
    "cfg2_k18": (100_000, 200, 18, 1e-3, 20240416, False),
    "cfg2_k19": (100_000, 200, 19, 1e-3, 20240416, False),
    "k20_m200": (100_000, 200, 20, 1e-3, 20240420, False),
    "cfg3": (1_000_000, 512, 24, 0.0, 20240417, False),
    "m512_k16": (200_000, 512, 16, 0.0, 20240417, False),      # configs[2]'s width at a one-GPU orthant count
    "m512_k20": (200_000, 512, 20, 0.0, 20240417, False),
    "m512_k24": (200_000, 512, 24, 0.0, 20240417, False),      # one rank's share of configs[2] is solved as a range of this
    "small": (20_000, 64, 10, 1e-3, 20240415, False),
}


def make_synthetic(N, M, K, seed, mixed_sign=False, rho=0.0):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, M))
    if rho:
        for m in range(1, M):
            X[:, m] = rho * X[:, m - 1] + np.sqrt(1 - rho * rho) * X[:, m]
    g = (np.arange(M) * K) // M
    s = rng.choice([-1.0, 1.0], size=K)
    z = rng.standard_normal(M)
    w = z if mixed_sign else np.abs(z) * s[g]
    y = X @ w + 0.5 + rng.standard_normal(N)
    P = np.zeros((M, K), dtype=np.int64)
    P[np.arange(M), g] = 1
    return np.asfortranarray(X), y, np.asfortranarray(P)


def make_config(name):
    N, M, K, eta, seed, mixed = CONFIGS[name]
    X, y, P = make_synthetic(N, M, K, seed, mixed_sign=mixed)
    return X, y, P, eta


def make_synthetic_parallel(N, M, K, seed, nthreads=None):
    """Same model as make_synthetic for the multi-GB shapes (configs[2]: 4.1 GB): the columns of X are drawn in
    blocks by worker threads from independent child streams of `seed` (numpy's generators release the GIL), so
    the values differ from make_synthetic's single stream but are fixed for a given (seed, M) -- block
    boundaries do not depend on the thread count."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    nthreads = nthreads or min(16, os.cpu_count() or 1)
    X = np.empty((N, M), order="F")
    blk = 16
    starts = list(range(0, M, blk))
    children = np.random.SeedSequence(seed).spawn(len(starts) + 1)

    def fill(i):
        g = np.random.default_rng(children[i])
        for m in range(starts[i], min(M, starts[i] + blk)):
            g.standard_normal(N, out=X[:, m])

    with ThreadPoolExecutor(nthreads) as ex:
        list(ex.map(fill, range(len(starts))))
    rng = np.random.default_rng(children[-1])
    g = (np.arange(M) * K) // M
    s = rng.choice([-1.0, 1.0], size=K)
    w = np.abs(rng.standard_normal(M)) * s[g]
    y = X @ w + 0.5 + rng.standard_normal(N)
    P = np.zeros((M, K), dtype=np.int64)
    P[np.arange(M), g] = 1
    return X, y, np.asfortranarray(P)
