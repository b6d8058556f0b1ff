"""Numpy model of the two-level K2 solver (design-time tool, not shipped, not the oracle).

State: a symmetric tableau T = sweep([G c; c' yy], O) over a set O of "stable" passive variables,
plus a small explicit inverse Sig = inv(S[FI, FI]) over the passive variables FI that are NOT swept
(S = Schur complement = T restricted to un-swept variables).  Block principal pivoting runs on the
reduced problem (T, c~ = T[:, rhs]); swept variables are checked through their implied weights
T[m, rhs] - T[m, FI] w_FI.  Counts how often the expensive full-tableau operations happen.

  python tools/proto_twolevel.py N M K L_inner n_steps [rho]
"""
import sys, time
import numpy as np
sys.path.insert(0, "tools")
from proto_bpp import make_cfg, gram, gray


class TwoLevel:
    def __init__(self, G, c, yy, gmask, l_inner):
        self.Mp = len(c)
        Mp = self.Mp
        self.T = np.zeros((Mp + 1, Mp + 1))
        self.T[:Mp, :Mp] = G; self.T[:Mp, Mp] = c; self.T[Mp, :Mp] = c; self.T[Mp, Mp] = yy
        self.G, self.c = G, c
        self.swp = np.zeros(Mp, bool)
        self.F = np.zeros(Mp, bool)          # passive and un-swept (in Sig)
        self.gmask = gmask
        self.low = (1 << l_inner) - 1
        self.nocommit = np.zeros(Mp, bool)
        self.n_sweep_blocks = self.n_unsweep_blocks = self.n_fickle = 0
        self.n_iter = self.n_inner_piv = 0
        self.scale = np.abs(c).max()

    def sweep_in(self, B):
        T = self.T
        P = T[:, B].copy(); D = T[np.ix_(B, B)]; Di = np.linalg.inv(D)
        T -= P @ Di @ P.T
        T[:, B] = P @ Di; T[B, :] = (P @ Di).T
        T[np.ix_(B, B)] = -Di
        self.swp[B] = True; self.F[B] = False
        self.n_sweep_blocks += 1

    def unsweep(self, B):
        T = self.T
        P = T[:, B].copy(); D = T[np.ix_(B, B)]; Di = np.linalg.inv(D)
        T -= P @ Di @ P.T
        T[:, B] = -P @ Di; T[B, :] = -(P @ Di).T
        T[np.ix_(B, B)] = -Di
        self.swp[B] = False; self.F[B] = True
        self.n_unsweep_blocks += 1

    def solve(self, sigma):
        Mp = self.Mp; T = self.T
        t_best, pbar = Mp + 1, 3
        while True:
            self.n_iter += 1
            idx = np.flatnonzero(self.F)
            w = np.zeros(Mp)
            if len(idx):
                w[idx] = np.linalg.solve(T[np.ix_(idx, idx)], T[idx, Mp])
            r = T[:Mp, Mp] - T[:Mp, :Mp][:, idx] @ w[idx]
            # swept variables: r[m] is the implied weight
            viol_sw = self.swp & ((sigma * r < 0) | ((sigma == 0) & (r != 0)))
            if viol_sw.any():
                v = np.flatnonzero(viol_sw)
                for q in range(0, len(v), 8):
                    self.unsweep(v[q:q + 8])
                self.last_unswept = getattr(self, "last_unswept", []) + list(v)
                continue
            Vp = self.F & ((sigma * w < 0) | (sigma == 0))
            Vd = (~self.F) & (~self.swp) & (sigma != 0) & (sigma * r > 1e-12 * self.scale)
            V = Vp | Vd
            nv = int(V.sum())
            if nv == 0:
                wfull = w.copy(); wfull[self.swp] = r[self.swp]
                obj2 = T[Mp, Mp] - T[idx, Mp] @ w[idx]
                return wfull, obj2
            if nv < t_best: t_best, pbar = nv, 3
            elif pbar >= 1: pbar -= 1
            else:
                j = np.flatnonzero(V).max(); V = np.zeros(Mp, bool); V[j] = True
            self.n_inner_piv += int(V.sum())
            self.F ^= V

    def commit(self):
        el = self.F & ((self.gmask & self.low) == 0) & (~self.nocommit)
        v = np.flatnonzero(el)
        for q in range(0, len(v), 8):
            self.sweep_in(v[q:q + 8])


if __name__ == "__main__":
    N, M, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    l_inner = int(sys.argv[4]); n_steps = int(sys.argv[5])
    rho = float(sys.argv[6]) if len(sys.argv) > 6 else 0.0
    X, y, P = make_cfg(N, M, K, 20240416)
    if rho:
        rng = np.random.default_rng(5)
        for m in range(1, M): X[:, m] = rho * X[:, m - 1] + np.sqrt(1 - rho * rho) * X[:, m]
        g = (np.arange(M) * K) // M
        s = rng.choice([-1.0, 1.0], size=K); z = rng.standard_normal(M)
        y = X @ (np.abs(z) * s[g]) + 0.5 + rng.standard_normal(N)
    G, c, yy, Po = gram(X, y, P, 1e-3)
    Mp, Kp = Po.shape
    gmask = np.array([sum(1 << k for k in range(Kp) if Po[m, k]) for m in range(Mp)], dtype=np.int64)
    tl = TwoLevel(G, c, yy, gmask, l_inner)
    i0 = 12345 * 7 % (2 ** Kp - n_steps)
    pis = []; t0 = time.time(); maxerr = 0.0
    prev_sigma = None
    for i in range(i0, i0 + n_steps):
        b = gray(i)
        beta = np.array([2 * ((b >> k) & 1) - 1 for k in range(Kp)], float)
        sigma = np.sign(Po @ beta)
        tl.last_unswept = []
        w, obj2 = tl.solve(sigma)
        # fickle = un-swept although its own sign class did not change
        if prev_sigma is not None:
            for m in tl.last_unswept:
                if sigma[m] == prev_sigma[m]:
                    tl.nocommit[m] = True; tl.n_fickle += 1
        prev_sigma = sigma
        pis.append(int(tl.F.sum()))
        tl.commit()
        if (i - i0) % 37 == 0:
            Fall = tl.F | tl.swp
            idx = np.flatnonzero(Fall)
            wd = np.zeros(Mp); wd[idx] = np.linalg.solve(G[np.ix_(idx, idx)], c[idx])
            r = c - G @ wd
            assert (sigma[idx] * wd[idx] >= -1e-12).all()
            act = ~Fall & (sigma != 0)
            assert (sigma[act] * r[act] <= 1e-9 * np.abs(c).max()).all()
            maxerr = max(maxerr, np.abs(wd - w).max() / np.abs(wd).max())
    n = n_steps
    print(f"steps {n}: iters/orthant {tl.n_iter/n:.2f} inner pivots/orthant {tl.n_inner_piv/n:.2f} "
          f"sweep-in blocks/orthant {tl.n_sweep_blocks/n:.3f} unsweep blocks/orthant {tl.n_unsweep_blocks/n:.3f} "
          f"fickle {tl.n_fickle} nocommit {int(tl.nocommit.sum())}  p_I mean {np.mean(pis):.1f} max {np.max(pis)} "
          f"swept {int(tl.swp.sum())}  max rel err vs direct {maxerr:.2e}  ({time.time()-t0:.1f}s)")
