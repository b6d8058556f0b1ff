"""Numpy model of the K2 solver (block principal pivoting in signed-weight space,
warm-started along a Gray-code chain of orthants).  Design-time tool only: counts
pivots/iterations so the CUDA kernel can be sized.  Not shipped, not the oracle."""
import numpy as np, sys, time
from scipy.optimize import nnls

def make_cfg(N, M, K, seed, eta=1e-3, mixed=False):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, M))
    g = (np.arange(M) * K) // M
    s = rng.choice([-1.0, 1.0], size=K)
    z = rng.standard_normal(M)
    w = z if mixed else np.abs(z) * s[g]
    y = X @ w + 0.5 + rng.standard_normal(N)
    P = np.zeros((M, K), dtype=np.int64); P[np.arange(M), g] = 1
    return X, y, P

def gram(X, y, P, eta):
    N, M = X.shape; K = P.shape[1]
    Xo = np.hstack([X, np.ones((N, 1))])
    Po = np.zeros((M + 1, K + 1)); Po[:M, :K] = P; Po[M, K] = 1
    G = Xo.T @ Xo + eta * (Po @ Po.T)
    c = Xo.T @ y
    return G, c, float(y @ y), Po

def bpp(G, c, sigma, F0, maxit=1000, tol=1e-10):
    """sigma in {-1,0,1}^M'; F0 boolean warm-start passive set.  Returns w, F, n_iter, n_exch"""
    Mp = len(c)
    F = F0.copy() & (sigma != 0)
    scale = np.abs(c).max()
    nit = nex = 0
    pbar, t = 3, Mp + 1
    while True:
        nit += 1
        w = np.zeros(Mp)
        idx = np.flatnonzero(F)
        if len(idx):
            w[idx] = np.linalg.solve(G[np.ix_(idx, idx)], c[idx])
        r = c - G @ w
        Vp = F & (sigma * w < 0)
        Vd = (~F) & (sigma != 0) & (sigma * r > tol * scale)
        V = Vp | Vd
        nv = int(V.sum())
        if nv == 0:
            return w, F, nit, nex
        if nv < t:
            t, pbar = nv, 3
        elif pbar >= 1:
            pbar -= 1
        else:
            j = np.flatnonzero(V).max()
            V = np.zeros(Mp, bool); V[j] = True
        nex += int(V.sum())
        F = F ^ V
        if nit > maxit:
            raise RuntimeError("no convergence")

def gray(i): return i ^ (i >> 1)

if __name__ == "__main__":
    N, M, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    L = int(sys.argv[4]) if len(sys.argv) > 4 else 64
    X, y, P = make_cfg(N, M, K, 20240416)
    G, c, yy, Po = gram(X, y, P, 1e-3)
    Mp, Kp = Po.shape
    rng = np.random.default_rng(0)
    tot_it = tot_ex = 0; ps = []; n = 0
    t0 = time.time()
    for chain in range(4):
        base = int(rng.integers(0, 2**Kp // L)) * L
        F = np.zeros(Mp, bool)
        for i in range(L):
            b = base + gray(i)
            beta = np.array([2 * ((b >> k) & 1) - 1 for k in range(Kp)], float)
            sigma = np.sign(Po @ beta)
            w, F, nit, nex = bpp(G, c, sigma, F)
            if i < 3 or i == L - 1:
                print(f"chain {chain} step {i} b={b} p={F.sum()} iters={nit} exch={nex}")
            if i > 0:
                tot_it += nit; tot_ex += nex; n += 1
            ps.append(F.sum())
            if i % 16 == 5:   # spot check vs L-H in data space (via Gram-cholesky surrogate)
                d = Po @ beta
                Lc = np.linalg.cholesky(G); A = (Lc.T * d); bb = np.linalg.solve(Lc, c)
                a, _ = nnls(A, bb, maxiter=10 * Mp)
                err = np.abs(a * d - w).max() / np.abs(w).max()
                assert err < 1e-9, err
    print(f"warm steps: mean iters {tot_it/n:.2f} mean exch {tot_ex/n:.2f}; p mean {np.mean(ps):.1f} max {np.max(ps)}  ({time.time()-t0:.1f}s)")
