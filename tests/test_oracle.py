"""CPU tests: the oracle (numpy/scipy + C restatements) against the reference's known answers,
the committed golden vectors, and each other."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = ["rand_a", "rand_b", "rand_c_corr", "rand_d_overlap"]


def test_toy_reference_assertions(oracle):
    """test/runtests.jl:8-39 -- opt ~ 0 (atol 1e-6) and sum(y_pred - y)^2 ~ 0 for Opt, eta = 0."""
    o, oc = oracle
    for fit in (lambda: o.fit_opt(o.TOY_X, o.TOY_Y, o.TOY_P, 0.0),):
        r = fit()
        assert abs(r["opt"]) < 1e-6
        yp = o.predict(r["alpha"], r["beta"], r["t"], o.TOY_P, o.TOY_X)
        assert abs(np.sum(yp - o.TOY_Y) ** 2) < 1e-6
    c = oc.opt_fit(o.TOY_X, o.TOY_Y, o.TOY_P, 0.0)
    assert abs(c["obj_best"]) < 1e-6 and c["b_best"] == 5


def test_toy_known_answer(oracle):
    """b* = 5, alpha = [5/11, 6/11, 1], beta = [11/29, -16/29], t = 60/29."""
    o, oc = oracle
    g = np.load(os.path.join(GOLD, "toy.npz"))
    expect = np.array([3.872983346207417, 0.6158788124473322, 2.7723120830300267, 1.073849688342439,
                       0.7071067811865476, 0.0, 1.6583123951777, 1.1079149192658462])
    for r_objs, b, a_raw in ((g["objs"], int(g["b_best"]), g["alpha_raw"]),
                             (lambda c: (c["objs"], c["b_best"], c["alpha_best"]))(oc.opt_fit(o.TOY_X, o.TOY_Y, o.TOY_P, 0.0))):
        assert np.allclose(r_objs, expect, rtol=1e-12, atol=1e-12)
        assert b == 5
        assert np.allclose(a_raw, [5 / 29, 6 / 29, 16 / 29, 60 / 29], rtol=1e-12)
        a, bb, t = o.cleanup_result_opt(a_raw, o.index_to_beta(b, 3), o.TOY_P)
        assert np.allclose(a, [5 / 11, 6 / 11, 1.0], rtol=1e-12)
        assert np.allclose(bb, [11 / 29, -16 / 29], rtol=1e-12)
        assert abs(t - 60 / 29) < 1e-12


def test_toy_eta_known_answer(oracle):
    o, oc = oracle
    r = o.fit_opt(o.TOY_X, o.TOY_Y, o.TOY_P, 1e-3)
    assert r["b_best"] == 5
    assert abs(r["opt"] - 0.06842847794863811) < 1e-12
    assert np.allclose(r["alpha"], [0.4513527805661314, 0.5486472194338686, 1.0], rtol=1e-10)
    assert np.allclose(r["beta"], [0.3847771682801378, -0.5490316741766817], rtol=1e-10)
    assert abs(r["t"] - 2.0462364266456983) < 1e-10
    c = oc.opt_fit(o.TOY_X, o.TOY_Y, o.TOY_P, 1e-3)
    assert c["b_best"] == 5 and abs(c["obj_best"] - r["opt"]) < 1e-12


def test_index_to_beta_bit_order(oracle):
    o, _ = oracle
    assert list(o.index_to_beta(0, 3)) == [-1, -1, -1]        # Opt.jl:4-20: b = 0 -> all -1
    assert list(o.index_to_beta(5, 3)) == [1, -1, 1]          # LSB first
    assert list(o.index_to_beta(4, 3)) == [-1, -1, 1]


def test_regularize_is_gram_ridge_on_groups(oracle):
    """regularizeProblem rows (PartitionedLS.jl:115-118) == + eta * Po*Po' in Gram space."""
    o, _ = oracle
    X, y, P = o.make_synthetic(50, 7, 3, 3)
    P[0, 1] = 1
    Xo, Po = o.homogeneous_coords(X, P)
    Xa, ya = o.regularize_problem(Xo, y, Po, 0.37)
    assert Xa.shape == (50 + 4, 8) and ya.shape == (54,)
    assert np.allclose(Xa.T @ Xa, Xo.T @ Xo + 0.37 * (Po @ Po.T), rtol=1e-13)
    assert np.allclose(Xa.T @ ya, Xo.T @ y) and np.isclose(ya @ ya, y @ y)
    Xn, yn = o.regularize_problem(Xo, y, Po, 0.0)
    assert Xn is Xo and np.array_equal(yn, y)


@pytest.mark.parametrize("name", CASES)
def test_c_oracle_matches_golden(oracle, name):
    o, oc = oracle
    g = np.load(os.path.join(GOLD, name + ".npz"))
    c = oc.opt_fit(g["X"], g["y"], g["P"], float(g["eta"]))
    assert c["b_best"] == int(g["b_best"])
    assert np.allclose(c["objs"], g["objs"], rtol=1e-11, atol=1e-11)
    assert np.allclose(c["alphas"], g["alphas"], rtol=1e-9, atol=1e-11)
    a, b, t = o.cleanup_result_opt(c["alpha_best"], o.index_to_beta(c["b_best"], g["P"].shape[1] + 1), g["P"])
    assert np.allclose(a, g["alpha"], rtol=1e-9, atol=1e-12) and np.allclose(b, g["beta"], rtol=1e-9, atol=1e-12)
    assert abs(t - float(g["t"])) <= 1e-9 * max(1.0, abs(float(g["t"])))


def test_c_nnls_matches_scipy(oracle):
    o, oc = oracle
    from scipy.optimize import nnls
    rng = np.random.default_rng(0)
    for t in range(150):
        m, n = int(rng.integers(2, 40)), int(rng.integers(1, 30))
        A = rng.standard_normal((m, n)); b = rng.standard_normal(m)
        if t % 5 == 0 and n > 2:
            A[:, 1] = A[:, 0]                     # exactly dependent columns
        xs, rs = nnls(A, b)
        xc, rc, st = oc.nnls(A, b)
        assert st == 0 and np.all(xc >= 0)
        assert abs(rs - rc) <= 1e-9 * (1 + rs)


def test_c_oracle_sampled_orthants_and_threads(oracle):
    o, oc = oracle
    X, y, P = o.make_synthetic(300, 14, 4, 21, mixed_sign=True)
    full = oc.opt_fit(X, y, P, 1e-3, nthreads=1)
    bl = [3, 17, 0, 31, 8]
    part = oc.opt_fit(X, y, P, 1e-3, b_list=bl, nthreads=3)
    assert np.allclose(part["objs"], full["objs"][bl], rtol=1e-13)
    assert part["b_best"] == bl[int(np.argmin(full["objs"][bl]))]
    multi = oc.opt_fit(X, y, P, 1e-3, nthreads=4)
    assert np.array_equal(multi["objs"], full["objs"]) and multi["b_best"] == full["b_best"]


def test_bnb_and_alt_on_toy_and_random(oracle):
    """test/runtests.jl:41-69 (toy, all three algorithms) + BnB reaches the Opt optimum."""
    o, _ = oracle
    r = o.fit_bnb(o.TOY_X, o.TOY_Y, o.TOY_P)
    assert abs(r["opt"]) < 1e-6 and r["nopen"] == 1
    assert np.allclose(o.predict(r["alpha"], r["beta"], r["t"], o.TOY_P, o.TOY_X), o.TOY_Y, atol=1e-9)
    ra = o.fit_alt(o.TOY_X, o.TOY_Y, o.TOY_P, beta0=np.array([1.0, -2.0, 3.0]))
    assert abs(ra["opt"]) < 1e-6
    for seed in (1, 2, 3):
        X, y, P = o.make_synthetic(300, 10, 3, seed, mixed_sign=True)
        a, b = o.fit_opt(X, y, P, 0.0), o.fit_bnb(X, y, P, 0.0)
        assert abs(a["opt"] - b["opt"]) <= 1e-9 * a["opt"]
        assert b["nopen"] <= 2 ** (P.shape[1] + 2)
