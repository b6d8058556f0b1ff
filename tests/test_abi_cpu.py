"""CPU tests of the boundary: the shared library loads and exports every symbol include/pls.h
declares, struct layouts agree with the C header, the host mirror's own maths (cleanupResult,
predict, homogeneousCoords, regularizeProblem) matches the oracle, and the product path fails
loudly without a GPU (no CPU fallback).  No compute entry point is called here."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pls.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pls_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(pkg):
    names = _declared_functions()
    assert "pls_opt_fit" in names and "pls_create" in names and len(names) >= 15
    lib = ctypes.CDLL(pkg._abi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pls.h but not exported"


def test_version_matches_header(pkg):
    ver = int(re.search(r"#define PLS_VERSION (\d+)", open(HEADER).read()).group(1))
    assert pkg._abi.lib.pls_version() == ver


def test_stats_struct_layout_matches_header(pkg):
    code = '#include <stdio.h>\n#include <stddef.h>\n#include "pls.h"\nint main(void){printf("%zu %zu %zu\\n", sizeof(pls_stats), offsetof(pls_stats, orthants), offsetof(pls_stats, gram_flops));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c"); exe = os.path.join(d, "t")
        open(src, "w").write(code)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        size, off1, off2 = map(int, subprocess.check_output([exe]).split())
    S = pkg._abi.PlsStats
    assert ctypes.sizeof(S) == size and S.orthants.offset == off1 and S.gram_flops.offset == off2


def test_no_gpu_means_loud_failure(pkg):
    if pkg._abi.lib.pls_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(pkg.PlsError) as e:
        pkg.Context(0)
    assert e.value.code == pkg._abi.PLS_ECUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(pkg.PlsError):
        pkg.fit(pkg.Opt, np.ones((4, 3)), np.ones(4), np.array([[1, 0], [1, 0], [0, 1]]))


def test_product_does_not_import_oracle():
    """The product path must not route through oracle/ (checked on the sources)."""
    pdir = os.path.join(ROOT, "partitionedls.jl_b200")
    for dp, _, files in os.walk(pdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                txt = open(os.path.join(dp, f)).read()
                assert "pls_oracle" not in txt and "oracle_c" not in txt and "scipy" not in txt, f


def test_host_cleanup_predict_match_oracle(pkg, oracle):
    o, _ = oracle
    rng = np.random.default_rng(5)
    M, K = 9, 3
    P = np.zeros((M, K), dtype=np.int64); P[np.arange(M), np.arange(M) % K] = 1
    P[2, 0] = 1                                    # overlapping group
    araw = np.abs(rng.standard_normal(M + 1)); araw[[1, 4, 7]] = 0.0   # group 1 entirely zero
    for b in (0, 5, 11, 15):
        opt, model = pkg._cleanup_result(1.25, araw, b, P)
        a, bb, t = o.cleanup_result_opt(araw, o.index_to_beta(b, K + 1), P)
        assert opt == 1.25 and np.array_equal(model.α, a) and np.array_equal(model.β, bb) and model.t == t
        X = rng.standard_normal((6, M))
        assert np.allclose(pkg.predict(model, X), o.predict(a, bb, t, P, X), rtol=1e-14)
        assert np.allclose(pkg.predict(model.α, model.β, model.t, P, X), pkg.predict(model, X))


def test_host_rewrites_match_oracle(pkg, oracle):
    o, _ = oracle
    X, y, P = o.make_synthetic(20, 6, 2, 9)
    Xo, Po = pkg.homogeneousCoords(X, P)
    Xo2, Po2 = o.homogeneous_coords(X, P)
    assert np.array_equal(Xo, Xo2) and np.array_equal(Po, Po2)
    Xa, ya = pkg.regularizeProblem(Xo, y, Po, 0.5)
    Xa2, ya2 = o.regularize_problem(Xo2, y, Po2, 0.5)
    assert np.array_equal(Xa, Xa2) and np.array_equal(ya, ya2)


def test_no_cpu_fallback_for_any_algorithm(pkg, oracle):
    """Without a GPU every fit must fail loudly with PLS_ECUDA (no CPU fallback); unknown algorithm
    types are a TypeError.  Host-side pieces of BnB / Alt are checked against the oracle."""
    import ctypes
    o, _ = oracle
    X, y, P = np.ones((4, 3)), np.ones(4), np.array([[1, 0], [1, 0], [0, 1]])
    with pytest.raises(TypeError):
        pkg.fit(object, X, y, P)
    if pkg._abi.lib.pls_device_count() == 0:
        for alg in (pkg.Opt, pkg.BnB, pkg.Alt):
            with pytest.raises(pkg.PlsError) as e:
                pkg.fit(alg, X, y, P)
            assert e.value.code == pkg._abi.PLS_ECUDA
    # BnB post-processing (BnB.jl:36-39) and the Alt start-value stream (Alt.jl:65-66)
    a_s = np.array([0.5, 1.5, -2.0, 0.25])
    m = pkg._bnb_postprocess(a_s, P)
    assert np.allclose(m.β, [2.0, -2.0]) and np.allclose(m.α, [0.25, 0.75, 1.0]) and m.t == 0.25
    b0 = pkg.draw_alt_starts(3, 4, 3, restarts=2)
    g = np.random.default_rng(3); g.random(4); first = (g.random(3) - 0.5) * 10
    assert b0.shape == (3, 2) and np.array_equal(b0[:, 0], first) and np.all(np.abs(b0) <= 5)
    assert ctypes.sizeof(pkg._abi.PlsStats) == 8 * 27
