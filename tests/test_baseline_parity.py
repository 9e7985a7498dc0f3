"""Oracle parity AT THE SIZES BASELINE.json NAMES (round-1 verdict, "What's weak" 2): the headline grid, the adapt!
grid and the basis grid are compared with the CPU oracle itself, not with other engines of this library.

  cfg3  D = 10, 652 065 nodes: BLOCK walk, two points per thread, cell-key ordered batch >= 2^20 with edge / NaN /
        out-of-domain points mixed in; unmasked, depth-masked (3, 6, 8) and the fused surplus epilogue, against
        asgo_interpolate_batch (long-double sum, src/class.jl:629-680) with the abssum-scaled tolerance
  cfg5  D = 4, 46 721 nodes: basis(X, G) rows and basis(G) rows against asgo_basis_dense, pattern exact
        (src/class.jl:434-454, 483-538)
  cfg4  D = 6 train!(7, 6:11) + adapt!(13, 3e-3): final node SET against OracleGrid.adapt (src/train.jl:235-463)
"""
import numpy as np
import pytest

from conftest import assert_close, sinsum

pytestmark = pytest.mark.gpu


def _mirror(asg, OG):
    G = asg.SparseGridInterpolant(OG.D, lb=OG.lb, ub=OG.ub)
    G.levels, G.indices, G.depths = OG.levels.copy(), OG.indices.copy(), OG.depths.copy()
    G.fvals, G.coefs = OG.fvals.copy(), OG.coefs.copy()
    return G


@pytest.fixture(scope="module")
def cfg3(oracle):
    OG = oracle.OracleGrid(10)
    OG.rsg(8, 8)                                   # the oracle's own enumeration (src/train.jl:40-78)
    assert len(OG) == 652065
    rng = np.random.default_rng(1234)
    OG.coefs = rng.standard_normal(len(OG)) * 2.0 ** (-OG.depths.astype(np.float64))
    return OG


def _cfg3_batch(rng, M):
    D = 10
    X = rng.random((M, D))
    # edge material spread over the batch: knots of every level, faces, corners, outside, NaN, Inf, -0.0
    k = np.arange(0, M, 997)
    X[k] = rng.integers(0, 129, size=(len(k), D)) / 128.0
    k = np.arange(5, M, 4099)
    X[k] = rng.integers(0, 2, size=(len(k), D)).astype(np.float64)
    k = np.arange(11, M, 5003)
    X[k, rng.integers(0, D, size=len(k))] = 1.0
    X[17::7001, 0] = 1.25
    X[23::9001, 9] = -1e-9
    X[29::11003, 4] = np.nan
    X[31::13001, 7] = np.inf
    X[37::15013, 2] = -0.0
    return X


def test_cfg3_full_size_against_oracle(asg, cfg3):
    """The headline configuration at full size: 652 065 nodes, path = block, R = 2, ordered batch >= 2^20."""
    OG = cfg3
    G = _mirror(asg, OG)
    info = G.info()
    assert info["N"] == 652065 and info["subspaces"] == 19448 and info["block"] == 1
    G.set_path("block")
    rng = np.random.default_rng(77)
    M = (1 << 20) + 4321
    X = _cfg3_batch(rng, M)
    # sample: every kind of edge point + random interior points, 2 000 in all
    special = np.concatenate([np.arange(0, M, 997)[:150], np.arange(5, M, 4099)[:60], np.arange(11, M, 5003)[:60],
                              np.arange(17, M, 7001)[:30], np.arange(23, M, 9001)[:30], np.arange(29, M, 11003)[:30],
                              np.arange(31, M, 13001)[:30], np.arange(37, M, 15013)[:30]])
    sel = np.unique(np.concatenate([special, rng.choice(M, size=2000 - len(special), replace=False)]))
    y = G.evaluate(X)
    want, sc = OG.evaluate(X[sel], return_abssum=True)
    assert_close(y[sel], want, sc, "cfg3 full size, block walk")
    with np.errstate(invalid="ignore"):
        outside = ~np.all((X >= 0.0) & (X <= 1.0), axis=1)
    assert outside.sum() > 300 and np.all(y[outside] == 0.0)       # exact zeros (src/math.jl:349-353)
    assert np.all(np.isfinite(y))
    # depth masks (_interpolate_until_level, src/class.jl:666-680) and the surplus epilogue (src/train.jl:176-183)
    f = rng.standard_normal(M)
    sub = sel[::4]
    G_masked = {}
    for L in (3, 6, 8):
        ym = G.evaluate(X, max_depth=L)
        G_masked[L] = ym
        wm, scm = OG.evaluate(X[sub], max_depth=L, return_abssum=True)
        assert_close(ym[sub], wm, scm, f"cfg3 full size, depth <= {L}")
        al = G.surplus(X, f, max_depth=L)
        assert_close(al[sub], f[sub] - wm, scm + np.abs(f[sub]), f"cfg3 full size, surplus depth <= {L}")
    assert np.array_equal(G.evaluate(X, max_depth=8), y)            # depth 8 is the whole grid
    # PAIR mode of the walk (neighbours of the cell-key order share their coefficient loads; default from 5 points per
    # 2-bit cell on, i.e. 5.2e6 points for D = 10): forced on here, where most pairs straddle a cell boundary and go to
    # the follow-up launch -- the sums do not depend on the mode, bit for bit
    import os
    os.environ["ASG_BLOCK_PAIR"] = "1e-9"
    try:
        yp = G.evaluate(X)
        assert np.array_equal(yp, y)
        for L in (3, 6):
            assert np.array_equal(G.evaluate(X, max_depth=L), G_masked[L])
        assert np.array_equal(G.surplus(X, f, max_depth=6), f - G_masked[6])
    finally:
        del os.environ["ASG_BLOCK_PAIR"]
    # one point per thread (small batch) and the other walks of the same layout give the same sums
    small = G.evaluate(X[sel])
    assert_close(small, want, sc, "cfg3 full size, small batch")
    assert np.max(np.abs(small - y[sel])) <= 1e-13
    G.set_path("stream")
    assert_close(G.evaluate(X[sel]), want, sc, "cfg3 full size, stream walk")


def test_cfg5_basis_against_oracle(asg, oracle):
    """basis(X, G) and basis(G) on the D = 4 depth-10 grid (46 721 nodes): rows against the oracle's basis row."""
    OG = oracle.OracleGrid(4)
    OG.rsg(10, 10)
    assert len(OG) == 46721
    rng = np.random.default_rng(5)
    OG.coefs = rng.standard_normal(len(OG))
    G = _mirror(asg, OG)
    # evaluation of this grid takes the BLOCK walk since round 2 (suffix budgets up to 9 = levels up to 10)
    assert G.info()["block"] == 1
    Xe = np.vstack([rng.random((4000, 4)), rng.integers(0, 1025, size=(500, 4)) / 1024.0])
    Xe[::97, 1] = 1.0 + 1e-12
    we, se = OG.evaluate(Xe, return_abssum=True)
    for path in ("block", "stream", "sweep"):
        G.set_path(path)
        assert_close(G.evaluate(Xe), we, se, f"cfg5 evaluation, {path}")
        assert_close(G.evaluate(Xe, max_depth=7), OG.evaluate(Xe, max_depth=7), se, f"cfg5 evaluation, depth <= 7, {path}")
    G.set_path("auto")
    big = rng.random((300000, 4))                                   # two points per thread, ordered batch
    yb = G.evaluate(big)
    assert_close(yb[:3000], OG.evaluate(big[:3000]), None, "cfg5 large batch")
    X = rng.random((200, 4))
    X[:20] = rng.integers(0, 513, size=(20, 4)) / 512.0             # knots: extra supporting nodes with value 0 drop out
    B = asg.basis(X, G).toarray()
    want = OG.basis(X)
    assert np.array_equal(B != 0.0, want != 0.0)
    assert np.max(np.abs(B - want)) <= 1e-15
    assert_close(B @ OG.coefs, OG.evaluate(X), np.abs(want) @ np.abs(OG.coefs), "B*c == G(x)")   # README.md:148-155
    Bd = asg.basis(X, G, dense=True)
    assert np.array_equal(Bd, B)
    # basis(G) = dehierarchization matrix: 200 node rows
    rows = np.sort(rng.choice(len(OG), size=200, replace=False))
    E = asg.basis(G).tocsr()
    assert E.shape == (46721, 46721)
    Er = E[rows].toarray()
    wantE = OG.basis(OG.stack()[rows])
    assert np.array_equal(Er != 0.0, wantE != 0.0)
    assert np.max(np.abs(Er - wantE)) <= 1e-15
    assert np.all(E.diagonal() == 1.0)


def test_cfg4_adapt_node_set_against_oracle(asg, oracle):
    """train!(7, 6:11) + adapt!(13, tol 3e-3) on D = 6: the final node SET, fvals and coefficients against
    OracleGrid.adapt (rows compared as sorted sets)."""
    OG = oracle.OracleGrid(6)
    OG.train(sinsum, 7, tuple(range(6, 12)))
    olog = []
    OG.adapt(sinsum, 13, tol=3e-3, log=olog)
    G = asg.SparseGridInterpolant(6)
    asg.train(G, sinsum, 7, tuple(range(6, 12)))
    glog = []
    asg.adapt(G, sinsum, 13, tol=3e-3, verbose=False, log=glog)
    assert len(OG) == len(G) == 19929
    assert [e["added"] for e in olog] == [e["added"] for e in glog]
    assert min(e["min_margin"] for e in glog) > 1e-9                # no acceptance decision near the tolerance
    ko = np.lexsort(np.hstack([OG.levels, OG.indices]).T[::-1])
    kg = np.lexsort(np.hstack([G.levels, G.indices]).T[::-1])
    assert np.array_equal(OG.levels[ko], G.levels[kg]) and np.array_equal(OG.indices[ko], G.indices[kg])
    assert np.array_equal(OG.depths[ko], G.depths[kg])
    assert np.max(np.abs(OG.fvals[ko] - G.fvals[kg])) <= 1e-15
    assert np.max(np.abs(OG.coefs[ko] - G.coefs[kg])) <= 1e-13
    # and the grown grid evaluates like the oracle's
    rng = np.random.default_rng(9)
    X = rng.random((3000, 6))
    want, sc = OG.evaluate(X, return_abssum=True)
    assert_close(G.evaluate(X), want, sc, "cfg4 after adapt!")
