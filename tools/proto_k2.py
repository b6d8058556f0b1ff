"""Line-by-line numpy model of the K2 CUDA kernel (csrc/nnls.cu): same state, same update
formulas, same pivoting rules.  Design-time check of the algorithm on the CPU; not shipped."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import pls_oracle as o

class Chain:
    def __init__(s, G, c, yy, gmask):
        s.G, s.c, s.yy, s.gmask = G, c, yy, gmask
        s.Mp = len(c); s.H = np.zeros((s.Mp, s.Mp)); s.p = 0
        s.w = np.zeros(s.Mp); s.r = c.copy(); s.pos = -np.ones(s.Mp, int); s.F = np.zeros(s.Mp, int)
        s.cmax = np.abs(c).max(); s.npiv = 0; s.ngrad = 0; s.nblocked = 0
    def add(s, j):
        p = s.p; F = s.F[:p]
        v = s.G[F, j]; u = s.H[:p, :p] @ v
        a = v @ u; b = v @ s.w[F]
        gjj = s.G[j, j]; delta = gjj - a
        if not (delta > 1e-13 * gjj): s.nblocked += 1; return False
        inv = 1 / delta; theta = (s.c[j] - b) * inv
        s.H[:p, :p] += np.outer(u * inv, u)
        s.H[p, :p] = -u * inv; s.H[:p, p] = -u * inv; s.H[p, p] = inv
        s.w[F] -= theta * u; s.w[j] = theta; s.F[p] = j; s.pos[j] = p; s.p += 1; s.npiv += 1
        return True
    def remove(s, slot):
        p = s.p; u = s.H[:p, slot].copy(); hss = u[slot]; j = s.F[slot]; inv = 1 / hss; f = s.w[j] * inv
        s.H[:p, :p] -= np.outer(u * inv, u)
        for t in range(p):
            if t != slot: s.w[s.F[t]] -= u[t] * f
        last = p - 1
        if slot != last:
            row = s.H[last, :last].copy(); corner = s.H[last, last]
            s.H[slot, :last] = row; s.H[:last, slot] = row; s.H[slot, slot] = corner
        s.w[j] = 0; s.pos[j] = -1
        if slot != last: jl = s.F[last]; s.F[slot] = jl; s.pos[jl] = slot
        s.p = last; s.npiv += 1
    def grad(s):
        F = s.F[:s.p]; s.r = s.c - s.G[:, F] @ s.w[F]; s.ngrad += 1
    def refine(s):
        F = s.F[:s.p]; v = s.r[F]; mx = np.abs(v).max() if s.p else 0.0
        if s.p: s.w[F] += s.H[:s.p, :s.p] @ v
        return mx
    def solve(s, b, r_valid):
        Mp = s.Mp
        d = np.array([2 * bin(int(g) & b).count("1") - bin(int(g)).count("1") for g in s.gmask])
        sg = np.sign(d); vflag = np.zeros(Mp, int)
        told = 1e-12 * s.cmax; t_best, pbar, iters = Mp + 1, 3, 0
        while True:
            if not r_valid:
                rep = 0
                while True:
                    s.grad(); rf = s.refine()
                    if rf <= 1e-9 * s.cmax: break
                    rep += 1
                    assert rep < 4, "refinement failed"
            r_valid = False
            rm, ad = [], []
            for m in range(Mp):
                if s.pos[m] >= 0:
                    if sg[m] == 0 or sg[m] * s.w[m] < 0: rm.append(m)
                elif sg[m] != 0 and vflag[m] != 3 and sg[m] * s.r[m] > told: ad.append(m)
            nv = len(rm) + len(ad)
            if nv == 0: break
            single = False
            if nv < t_best: t_best, pbar = nv, 3
            elif pbar >= 1: pbar -= 1
            else: single = True
            if single:
                m = max(rm + ad)
                if s.pos[m] >= 0: s.remove(s.pos[m])
                elif not s.add(m): vflag[m] = 3
            else:
                for m in rm: s.remove(s.pos[m])
                for m in ad:
                    if not s.add(m): vflag[m] = 3
            iters += 1
            assert iters < 60 + 6 * Mp
        F = s.F[:s.p]
        obj = np.sqrt(max(s.yy - s.c[F] @ s.w[F], 0.0))
        alpha = np.zeros(Mp)
        for m in F:
            if d[m] != 0: alpha[m] = max(s.w[m] / d[m], 0.0)
        return obj, alpha

def run(X, y, P, eta, L):
    Xo, Po = o.homogeneous_coords(X, P)
    G = Xo.T @ Xo + eta * (Po @ Po.T); c = Xo.T @ y; yy = y @ y
    Mp, Kp = Po.shape
    gmask = [sum(1 << k for k in range(Kp) if Po[m, k]) for m in range(Mp)]
    objs = np.zeros(2 ** Kp); alphas = np.zeros((2 ** Kp, Mp))
    for chain in range(2 ** Kp // L):
        ch = Chain(G, c, yy, gmask); rv = True
        for i in range(L):
            b = chain * L + (i ^ (i >> 1))
            objs[b], alphas[b] = ch.solve(b, rv); rv = True
    return objs, alphas

if __name__ == "__main__":
    worst = 0
    for seed in range(6):
        N, M, K = [(300, 14, 4), (500, 20, 5), (64, 9, 3), (2000, 30, 6), (40, 12, 4), (1000, 24, 3)][seed]
        X, y, P = o.make_synthetic(N, M, K, seed, mixed_sign=(seed % 2 == 1), rho=0.6 if seed >= 3 else 0.0)
        if seed == 4:  # overlapping groups + a feature in no group
            P[0, 1] = 1; P[3, :] = 0; P[5, 2] = 1
        for eta in (0.0, 1e-2):
            ref = o.fit_opt(X, y, P, eta, return_all=True)
            for L in (1, 8, 2 ** (K + 1)):
                objs, alphas = run(X, y, P, eta, L)
                eo = np.abs(objs - ref["objs"]).max() / ref["objs"].max()
                # compare alphas orthant by orthant
                ea = 0.0
                Xo, Po = o.homogeneous_coords(X, P); Xa, ya = o.regularize_problem(Xo, y, Po, eta)
                for b in range(len(objs)):
                    _, a_ref, _ = o.opt_orthant(Xa, ya, Po, b)
                    ea = max(ea, np.abs(alphas[b] - a_ref).max() / max(1e-300, np.abs(a_ref).max()))
                worst = max(worst, eo, ea)
                print(f"seed {seed} eta {eta} L {L}: obj rel err {eo:.2e} alpha rel err {ea:.2e} argmin {int(np.argmin(objs))} vs {ref['b_best']}")
                assert int(np.argmin(objs)) == ref["b_best"]
    # toy
    objs, alphas = run(o.TOY_X, o.TOY_Y, o.TOY_P, 0.0, 4)
    print(objs, alphas[5])
    print("worst", worst)
