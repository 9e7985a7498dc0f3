"""CPU tests: pin the oracle. The reference ships no tests; the known answers below are the ones its
docstrings / README state, plus analytic vectors (tests/golden/) that any
correct implementation must reproduce, plus agreement between the two independent restatements."""
import math
import os

import numpy as np
import pytest

from conftest import assert_close, parabola, sinsum, utility

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_docstring_known_answers(oracle):
    o = oracle
    # src/math.jl:513-514
    assert [o.fraction2li(i, 4) for i in range(5)] == [(2, 0), (3, 1), (1, 1), (3, 3), (2, 2)]
    # src/math.jl:626-627
    assert o.allnodes1d(1) == ([1], [1]) and o.allnodes1d(2) == ([1, 2, 2], [1, 0, 2])
    Ls, Is = o.allnodes1d(4)
    assert [o.get_x1(l, i) for l, i in zip(Ls, Is)] == [k / 8 for k in range(9)]
    # src/math.jl:651-661
    assert o.allindices_of_level(1).tolist() == [1] and o.allindices_of_level(2).tolist() == [0, 2]
    assert o.allindices_of_level(5).tolist() == list(range(1, 16, 2))
    with pytest.raises(o.ArgumentError):
        o.allindices_of_level(0)
    # src/math.jl:997-999, 1008-1019
    assert o.parent1(3, 1) == (2, 0) and o.parent1(3, 3) == (2, 2)
    assert o.parent1(2, 0) == (1, 1) and o.parent1(2, 2) == (1, 1)
    assert o.parent1(5, 11) == (4, 5) and o.parent1(4, 7) == (3, 3)
    # src/math.jl:59-61, 84-89
    assert [o.power2(k) for k in range(5)] == [1, 2, 4, 8, 16]
    # src/math.jl:907-959
    assert o.child1(1, 1, "left") == (2, 0) and o.child1(1, 1, "right") == (2, 2)
    assert o.child1(2, 0, "left") == (3, -1) and o.child1(2, 0, "right") == (3, 1)
    assert o.child1(2, 2, "left") == (3, 3) and o.child1(2, 2, "right") == (3, -1)
    assert o.child1(4, 3, "left") == (5, 5) and o.child1(4, 3, "right") == (5, 7)
    with pytest.raises(o.ArgumentError):
        o.child1(2, 1, "left")
    # src/math.jl:411-416
    assert o.isvalid([1, 2, 3], [1, 0, 3]) and not o.isvalid([1, 2, 3], [1, 1, 3])
    assert not o.isvalid([1, 3], [1, 2]) and not o.isvalid([0], [1]) and not o.isvalid([1], [0])
    assert not o.isvalid([3], [-1]) and o.isvalid([3], [5])   # no upper bound check in the reference
    # src/math.jl:464-476 / 355
    assert o.get_x1(1, 1) == 0.5 and o.get_x1(2, 0) == 0.0 and o.get_x1(2, 2) == 1.0 and o.get_x1(4, 3) == 0.375
    for bad in [(1, 0), (2, 1), (3, 2), (0, 1)]:
        with pytest.raises(o.ArgumentError):
            o.get_x1(*bad)
    with pytest.raises(o.ArgumentError):
        o.phi(0.5, 2, 1)
    with pytest.raises(o.ArgumentError):
        o.phi(0.5, 0, 1)


def test_phi_edge_semantics(oracle):
    o = oracle
    assert o.phi(float("nan")) == 0.0 and o.phi(1.0) == 0.0 and o.phi(-1.0) == 0.0 and o.phi(0.0) == 1.0
    assert o.phi(0.25, 1, 1) == 1.0 and o.phi(1.0000001, 1, 1) == 0.0 and o.phi(float("nan"), 1, 7) == 0.0
    assert o.phi(0.25, 2, 0) == 0.5 and o.phi(0.75, 2, 0) == 0.0 and o.phi(0.75, 2, 2) == 0.5
    assert o.phi(0.5, 2, 0) == 0.0 and o.phi(0.5, 2, 2) == 0.0
    assert o.phi(0.3125, 4, 3) == 0.5 and o.phi(0.375, 4, 3) == 1.0 and o.phi(0.5, 4, 3) == 0.0
    assert o.phi(-0.1, 3, 1) == 0.0 and o.phi(2.0, 3, 5) == 0.0
    # isbig2add, src/math.jl:257-274
    assert o.isbig2add(2e-3, 5.0, float("nan"), 1e-3, False) and not o.isbig2add(1e-3, 5.0, float("nan"), 1e-3, False)
    assert o.isbig2add(0.2, 1.0, 0.1, float("nan"), True) and not o.isbig2add(0.05, 1.0, 0.1, float("nan"), True)
    assert o.isbig2add(0.2, 0.05, 0.1, float("nan"), True) and not o.isbig2add(0.05, 0.05, 0.1, float("nan"), True)


def test_phi_dyadic_golden(oracle):
    g = np.load(os.path.join(GOLD, "phi_dyadic.npz"))
    got = np.array([oracle.phi(float(x), int(l), int(i)) for x, l, i in zip(g["x"], g["l"], g["i"])])
    assert np.array_equal(got, g["phi"])
    from oracle import oracle_np as onp
    for x in np.unique(g["x"])[:50]:
        m = g["x"] == x
        assert np.array_equal(onp.phi_vec(float(x), g["l"][m], g["i"][m]), g["phi"][m])


def test_rsg_order_and_sizes(oracle):
    G = oracle.OracleGrid(3)
    G.rsg(7, 5)
    assert len(G) == 737
    assert np.bincount(G.depths).tolist() == [0, 1, 6, 18, 44, 108, 216, 344]
    assert G.levels[:8].tolist() == [[1, 1, 1], [2, 1, 1], [2, 1, 1], [3, 1, 1], [3, 1, 1], [4, 1, 1], [4, 1, 1], [4, 1, 1]]
    assert G.indices[:8].tolist() == [[1, 1, 1], [0, 1, 1], [2, 1, 1], [1, 1, 1], [3, 1, 1], [1, 1, 1], [3, 1, 1], [5, 1, 1]]
    # brute-force Iterators.product enumeration (src/train.jl:56-63) for a small case
    import itertools
    D, md, ml = 3, 5, (3, 4, 2)
    exp_L, exp_I = [], []
    for ks in itertools.product(*[range(1, m + 1) for m in ml][::-1]):
        ks = ks[::-1]                                     # first dimension fastest
        if sum(ks) <= md + D - 1:
            idx = [oracle.allindices_of_level(k).tolist() for k in ks]
            for Is in itertools.product(*idx[::-1]):
                exp_L.append(ks)
                exp_I.append(Is[::-1])
    G2 = oracle.OracleGrid(D)
    G2.rsg(md, ml)
    assert G2.levels.tolist() == [list(k) for k in exp_L] and G2.indices.tolist() == [list(k) for k in exp_I]
    G10 = oracle.OracleGrid(10)
    G10.rsg(8, 8)
    assert len(G10) == 652065 and np.bincount(G10.depths).tolist()[1:] == [1, 20, 200, 1360, 7220, 32464, 130160, 480640]
    with pytest.raises(oracle.ArgumentError):
        G2.rsg(1, 3)
    with pytest.raises(oracle.ArgumentError):
        G2.rsg(3, 0)


def test_analytic_surpluses(oracle):
    g = np.load(os.path.join(GOLD, "parabola_d3.npz"))
    G = oracle.OracleGrid(3)
    G.train(parabola, 6, 6)
    assert np.array_equal(G.levels, g["levels"]) and np.array_equal(G.indices, g["indices"])
    assert np.array_equal(G.coefs, g["coefs"])             # bit-exact: every intermediate is dyadic
    assert np.array_equal(G.evaluate(g["X"]), g["y"])      # exact rational evaluation, dyadic points
    # affine function: surpluses at depth >= 3 vanish, root = f(centre)
    A = oracle.OracleGrid(2, lb=[0, 0], ub=[1, 1])
    A.train(lambda x: 1.0 + 2.0 * x[0] - 0.5 * x[1], 5, 5)
    assert A.coefs[0] == 1.75 and np.all(A.coefs[A.levels.max(axis=1) >= 3] == 0.0)


def test_cfg1_snapshot_and_identities(oracle):
    g = np.load(os.path.join(GOLD, "cfg1_eval.npz"))
    G = oracle.OracleGrid(3)
    G.train(sinsum, 7, 5)
    assert np.array_equal(G.levels, g["levels"]) and np.array_equal(G.indices, g["indices"])
    assert_close(G.coefs, g["coefs"], None, "coefs")       # libm sin may differ in the last ulp across boxes
    y, ab = G.evaluate(g["X"], return_abssum=True)
    assert_close(y, g["y"], g["abssum"], "eval")
    assert_close(G.evaluate(g["X"], max_depth=4), g["y_depth4"], g["abssum"], "masked eval")
    # interpolation property (src/train.jl:176-183): G(x_n) == fvals[n]
    assert np.max(np.abs(G.evaluate(G.stack()) - G.fvals)) < 1e-15 * 8
    # README.md:148-167: sum(B .* C) == G(x); E*coefs = fvals; E unit lower-triangular under a depth sort
    X = g["X"][:50]
    B = G.basis(X)
    assert_close(B @ G.coefs, G.evaluate(X), None, "B*C")
    E = G.basis()
    assert E.shape == (737, 737)
    assert_close(E @ G.coefs, G.fvals, None, "E*c")
    order = np.argsort(G.depths, kind="stable")
    Es = E[np.ix_(order, order)]
    assert np.all(np.diag(Es) == 1.0) and np.all(np.triu(Es, 1) == 0.0)
    # out-of-domain / NaN -> exactly 0 (src/class.jl:624-625)
    bad = np.array([[1.5, 0.5, 0.5], [0.5, -1e-9, 0.5], [0.5, 0.5, np.nan], [np.inf, 0.2, 0.2]])
    assert np.all(G.evaluate(bad) == 0.0)


def test_cfg2_adapt_snapshot(oracle):
    g = np.load(os.path.join(GOLD, "cfg2_nodes.npz"))
    G = oracle.OracleGrid(2)
    G.train(utility, 4, (3, 8))
    assert len(G) == 25
    log = []
    G.adapt(utility, 20, tol=1e-5, log=log)
    # per-level counts equal SURVEY.md appendix B (an independent restatement): 28, 64, ... 397 -> 12 119
    assert [l["added"] for l in log] == g["added"].tolist() == [28, 64, 144, 319, 583, 878, 1066, 1330, 1556, 1666,
                                                                 1377, 1122, 726, 500, 338, 397]
    key = np.concatenate([G.levels, G.indices], axis=1)
    order = np.lexsort(key.T[::-1])
    assert np.array_equal(key[order], g["key"])
    assert_close(G.coefs[order], g["coefs"], None, "adapt coefs")
    assert G.levels.max(axis=0).tolist() == [14, 17]
    assert min(l["min_margin"] for l in log) > 1e-13


def test_c_oracle_vs_numpy_restatement(oracle):
    from oracle import oracle_np as onp
    rng = np.random.default_rng(0)
    D = 4
    G = oracle.OracleGrid(D, lb=[-1, 0, 2, 10], ub=[1, 0.5, 7, 10.25])
    G.train(lambda x: float(np.sin(x[0]) * x[1] + np.cos(x[2]) - x[3] ** 2 / 100), 5, (3, 4, 5, 2))
    X = G.lb + (G.ub - G.lb) * rng.random((40, D))
    X[3, 1] = G.ub[1]
    X[4, 0] = G.lb[0]
    X[5] = G.lb + (G.ub - G.lb) * 0.25
    y = G.evaluate(X)
    ref = np.array([onp.interpolate(G.levels, G.indices, G.coefs, G.lb, G.ub, x) for x in X])
    assert np.max(np.abs(y - ref)) <= 2 * np.finfo(float).eps * np.max(np.abs(ref))   # long double vs exact fsum
    ym = G.evaluate(X, max_depth=3)
    refm = np.array([onp.interpolate(G.levels, G.indices, G.coefs, G.lb, G.ub, x, G.depths <= 3) for x in X])
    assert np.max(np.abs(ym - refm)) <= 2 * np.finfo(float).eps * np.max(np.abs(refm))
    Bm = G.basis(X[:5])
    for k in range(5):
        assert np.array_equal(Bm[k], onp.basis_row(G.levels, G.indices, G.lb, G.ub, X[k]))
    # the reference-structured baseline path (copy coefs, D passes, pairwise sum) agrees too
    assert np.max(np.abs(oracle.refpath_eval(G, X) - y)) < 1e-14


def test_basis_roundtrip_nonunit_domain(oracle):
    """A14 of SURVEY.md: basis(G) goes x01 -> x -> x01; the oracle follows the same round trip."""
    G = oracle.OracleGrid(2, lb=[1, 2], ub=[2, 3.3])
    G.train(lambda x: float(x[0] * x[1]), 5, 5)
    E = G.basis()
    assert_close(E @ G.coefs, G.fvals, None, "E*c non-unit")
    assert np.all(np.diag(E) > 0.999999)
