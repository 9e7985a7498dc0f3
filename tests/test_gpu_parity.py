"""GPU parity: libasg_cuda.so (through the C ABI) against the CPU oracle on identical inputs.

Bar (BASELINE.json north_star): node sets / levels bit-exact; fp64 values within 1e-12 relative,
1e-14 absolute near zero (conftest.assert_close). Sizes are chosen so the oracle finishes in seconds.
"""
import numpy as np
import pytest

from conftest import assert_close, parabola, sinsum, utility

pytestmark = pytest.mark.gpu


def _edge_points(D, lb, ub, rng, n=64):
    """knots, corners, faces, out-of-domain, NaN, +-0.0"""
    lb, ub = np.asarray(lb, float), np.asarray(ub, float)
    pts = []
    for k in range(0, 9):
        pts.append(lb + (ub - lb) * (k / 8.0))                  # diagonal knots incl. both corners
    dy = rng.integers(0, 65, size=(n, D)) / 64.0                  # dyadic points: exactly on knots
    pts.extend(lb + (ub - lb) * dy)
    c = rng.integers(0, 2, size=(16, D)).astype(float)            # corners
    pts.extend(lb + (ub - lb) * c)
    out = rng.random((8, D))
    out[:, 0] = 1.5                                               # outside in one dim
    pts.extend(lb + (ub - lb) * out)
    out2 = rng.random((8, D))
    out2[:, -1] = -0.25
    pts.extend(lb + (ub - lb) * out2)
    nanp = lb + (ub - lb) * rng.random((4, D))
    nanp[:, D // 2] = np.nan
    pts.extend(nanp)
    infp = lb + (ub - lb) * rng.random((2, D))
    infp[0, 0] = np.inf
    infp[1, 0] = -np.inf
    pts.extend(infp)
    z = np.array(lb, copy=True)
    if lb[0] == 0.0:
        z[0] = -0.0
    pts.append(z)
    return np.array(pts, dtype=np.float64)


def _mirror(asg, OG):
    G = asg.SparseGridInterpolant(OG.D, lb=OG.lb, ub=OG.ub)
    G.levels, G.indices, G.depths = OG.levels.copy(), OG.indices.copy(), OG.depths.copy()
    G.fvals, G.coefs = OG.fvals.copy(), OG.coefs.copy()
    return G


@pytest.fixture(scope="module")
def cfg1(oracle):
    OG = oracle.OracleGrid(3)
    OG.train(sinsum, 7, 5)
    return OG


@pytest.fixture(scope="module")
def cfg2(oracle):
    OG = oracle.OracleGrid(2)
    OG.train(utility, 4, (3, 8))
    OG.adapt(utility, 20, tol=1e-5)
    return OG


@pytest.fixture(scope="module")
def cfg_d10(oracle):
    OG = oracle.OracleGrid(10)
    OG.rsg(5, 5)
    rng = np.random.default_rng(1234)
    OG.coefs = rng.standard_normal(len(OG)) * 2.0 ** (-OG.depths.astype(float))
    return OG


@pytest.fixture(scope="module")
def cfg_nonunit(oracle):
    D = 4
    OG = oracle.OracleGrid(D, lb=np.arange(1, D + 1.0), ub=np.arange(2, D + 2.0) + 0.3)
    OG.train(lambda x: float(np.sum(np.cos(x)) + x[0] * x[1]), 6, (4, 5, 6, 3))
    return OG


@pytest.mark.parametrize("path", ["subspace", "trie", "sweep"])
@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg_d10", "cfg_nonunit"])
def test_eval_parity(asg, request, name, path):
    OG = request.getfixturevalue(name)
    G = _mirror(asg, OG)
    G.set_path(path)
    rng = np.random.default_rng(42)
    lb, ub = OG.lb, OG.ub
    X = np.vstack([lb + (ub - lb) * rng.random((4000, OG.D)), _edge_points(OG.D, lb, ub, rng)])
    want, scale = OG.evaluate(X, return_abssum=True)
    got = G.evaluate(X)
    assert_close(got, want, scale, f"{name}/{path}")
    # out-of-domain and NaN points are exactly 0.0 (src/math.jl:349-353)
    outside = ~np.all((X >= lb) & (X <= ub), axis=1)
    assert outside.sum() >= 20
    assert np.all(got[outside] == 0.0)
    info = G.info()
    assert info["path"] == (1 if path == "sweep" else 2)
    assert info["regular"] == (0 if name == "cfg2" else 1)    # regular grids take the record-free walk


def test_eval_layouts_and_scalar_call(asg, cfg1):
    G = _mirror(asg, cfg1)
    rng = np.random.default_rng(7)
    X = rng.random((1000, 3))
    want = cfg1.evaluate(X)
    a = G.evaluate(X)
    b = G.evaluate(np.asfortranarray(X))              # Julia column-major M x D
    c = G.evaluate(np.ascontiguousarray(X.T), dims=2)   # D x n as rand(G, n) returns
    d = G.evaluate(np.repeat(X, 2, axis=1)[:, ::2])    # generic strides
    assert_close(a, want, what="row-major")
    assert np.array_equal(a, b) and np.array_equal(a, c) and np.array_equal(a, d)
    assert G(X[0]) == a[0]
    assert np.array_equal(G(X[:5]), a[:5])
    with pytest.raises(AssertionError):
        G(np.zeros(4))
    with pytest.raises(NameError):
        G(X[0], extrapolate=True)                       # reference defect B3, mirrored


def test_interpolation_property(asg, cfg1, cfg_nonunit):
    """G(x_n) == fvals[n] at every node (residual fitting, src/train.jl:176-183)."""
    for OG in (cfg1, cfg_nonunit):
        G = _mirror(asg, OG)
        X = asg.stack(G)
        assert np.array_equal(X, OG.stack())            # device get_x is bit-exact (A2: no FMA)
        got = G.evaluate(X)
        assert np.max(np.abs(got - OG.fvals)) < 5e-14 * max(1.0, np.max(np.abs(OG.fvals)))


@pytest.mark.parametrize("path", ["subspace", "trie", "sweep"])
def test_masked_eval_and_surplus(asg, cfg1, cfg2, path):
    rng = np.random.default_rng(3)
    for OG in (cfg1, cfg2):
        G = _mirror(asg, OG).set_path(path)
        X = rng.random((500, OG.D))
        for L in (0, 1, 2, 3, 5, int(OG.depths.max()), int(OG.depths.max()) + 3):
            want, sc = OG.evaluate(X, max_depth=L, return_abssum=True)
            assert_close(G.evaluate(X, max_depth=L), want, sc, f"mask {L}")
        f = rng.standard_normal(500)
        want = f - OG.evaluate(X, max_depth=4)
        assert_close(G.surplus(X, f, max_depth=4), want, np.abs(f) + 1, "surplus")


def test_train_parity(asg, oracle):
    """train! on the GPU reproduces the oracle's grid: nodes bit-exact, coefficients in tolerance."""
    for D, f, depth, ml, lb, ub in [(3, sinsum, 7, 5, None, None), (2, utility, 4, (3, 8), None, None),
                                    (6, sinsum, 5, tuple(range(6, 12)), None, None),
                                    (3, lambda x: float(np.sum(np.sin(x))), 6, 6, [1, 2, 3], [2, 3, 4])]:
        OG = oracle.OracleGrid(D, lb=lb, ub=ub)
        OG.train(f, depth, ml)
        G = asg.SparseGridInterpolant(D, lb=lb, ub=ub)
        asg.train(G, f, depth, ml)
        assert np.array_equal(G.levels, OG.levels) and np.array_equal(G.indices, OG.indices)
        assert np.array_equal(G.depths, OG.depths)
        assert np.array_equal(G.fvals, OG.fvals)          # same f at bit-identical coordinates
        assert_close(G.coefs, OG.coefs, np.abs(OG.fvals) + 1, "train coefs")
        # device coefficients (written by the fused surplus kernel) equal the host copy
        import ctypes as C
        back = np.empty(len(G))
        asg._capi.check(asg._capi.load().asg_grid_get_coefs(G._ensure(), 0, len(G), back.ctypes.data_as(asg._capi.f64p)))
        assert np.array_equal(back, G.coefs)


def test_train_analytic_exact(asg):
    """f = prod 4x(1-x): every intermediate is dyadic, so surpluses are bit-exact."""
    G = asg.SparseGridInterpolant(3)
    asg.train(G, parabola, 6, 6)
    s1 = lambda l: 1.0 if l == 1 else (-1.0 if l == 2 else 4.0 ** (2 - l))
    exp = np.array([np.prod([s1(l) for l in L]) for L in G.levels])
    assert len(G) == 441 and np.array_equal(G.coefs, exp)
    # affine f: all surpluses at depth >= 3 are exactly 0, root = f(0.5,...)
    G2 = asg.SparseGridInterpolant(2)
    asg.train(G2, lambda x: 1.0 + 2.0 * x[0] - 0.5 * x[1], 5, 5)
    assert G2.coefs[0] == 1.0 + 1.0 - 0.25
    assert np.all(G2.coefs[G2.levels.max(axis=1) >= 3] == 0.0)


def test_adapt_parity(asg, oracle, cfg2):
    """adapt! node SET bit-exact against the oracle (row order is Dict order in Julia: compare sorted)."""
    G = asg.SparseGridInterpolant(2)
    asg.train(G, utility, 4, (3, 8))
    log = []
    asg.adapt(G, utility, 20, tol=1e-5, toltype="absolute", verbose=False, log=log)
    OG = cfg2
    assert len(G) == len(OG) == 12119

    def keyed(Lv, Ix, vals):
        key = np.concatenate([Lv, Ix], axis=1)
        order = np.lexsort(key.T[::-1])
        return key[order], vals[order]

    kg, cg = keyed(G.levels, G.indices, G.coefs)
    ko, co = keyed(OG.levels, OG.indices, OG.coefs)
    _, fo = keyed(OG.levels, OG.indices, OG.fvals)
    assert np.array_equal(kg, ko)
    # alpha = f - G inherits the ABSOLUTE error of G, whose terms are as large as |f| (up to 10 here)
    assert_close(cg, co, 4 * (np.abs(fo) + 1), "adapt coefs")
    assert min(l["min_margin"] for l in log) > 1e-13      # no acceptance decision within rounding of tol (H6)
    # and the adapted grid evaluates like the oracle's
    rng = np.random.default_rng(5)
    X = rng.random((2000, 2))
    want, sc = OG.evaluate(X, return_abssum=True)
    assert_close(G.evaluate(X), want, sc, "adapted eval")
    assert G.info()["hash_slots"] > 0                      # sparsely filled subspaces were exercised


def test_adapt_relative_d6(asg, oracle):
    f = sinsum
    OG = oracle.OracleGrid(4)
    OG.train(f, 4, (3, 4, 5, 6))
    OG.adapt(f, 7, tol=5e-2, toltype="relative")
    G = asg.SparseGridInterpolant(4)
    asg.train(G, f, 4, (3, 4, 5, 6))
    asg.adapt(G, f, 7, tol=5e-2, toltype="relative", verbose=False)
    a = {tuple(r) for r in np.concatenate([G.levels, G.indices], axis=1)}
    b = {tuple(r) for r in np.concatenate([OG.levels, OG.indices], axis=1)}
    assert a == b and len(G) > 100


@pytest.mark.parametrize("path", ["subspace", "sweep"])
def test_basis_parity(asg, cfg1, cfg2, cfg_nonunit, path):
    rng = np.random.default_rng(11)
    for OG in (cfg1, cfg2, cfg_nonunit):
        G = _mirror(asg, OG).set_path(path)
        lb, ub = OG.lb, OG.ub
        X = np.vstack([lb + (ub - lb) * rng.random((200, OG.D)), _edge_points(OG.D, lb, ub, rng, 16)])
        want = OG.basis(X)
        B = asg.basis(X, G)
        assert B.shape == (len(X), len(OG))
        got = B.toarray()
        assert np.array_equal(got != 0, want != 0)         # same sparsity pattern
        assert_close(got, want, None, "basis values")
        dense = asg.basis(X, G, dense=True)
        assert np.array_equal(dense, got)
        # README identity: sum(B .* C) == G(x)
        assert_close(B @ OG.coefs, OG.evaluate(X), None, "B*C")
        # dropzeros = false keeps tiny values, atol filter otherwise
        w2 = OG.basis(X, dropzeros=False)
        g2 = asg.basis(X, G, dropzeros=False).toarray()
        assert np.array_equal(g2 != 0, w2 != 0)
        w3 = OG.basis(X, atol=1e-3)
        g3 = asg.basis(X, G, atol=1e-3).toarray()
        assert np.array_equal(g3 != 0, w3 != 0)
        # single point form
        b1 = asg.basis(X[0], G)
        assert b1.shape == (1, len(OG)) and np.array_equal(b1.toarray()[0], got[0])


def test_dehierarchization(asg, cfg1):
    G = _mirror(asg, cfg1)
    E = asg.dehierarchization_matrix(G)
    assert E.shape == (len(G), len(G))                    # README.md:162-163
    assert np.array_equal(E.toarray(), cfg1.basis())
    assert_close(E @ G.coefs, G.fvals, None, "E*coefs = fvals")
    order = np.argsort(G.depths, kind="stable")
    Es = E.toarray()[np.ix_(order, order)]
    assert np.all(np.diag(Es) == 1.0) and np.all(np.triu(Es, 1) == 0.0)   # unit lower-triangular under depth sort
    H = asg.hierarchization_matrix(G)
    assert_close(H @ G.fvals, G.coefs, np.abs(G.fvals).max() * np.ones(len(G)) * 10, "H*fvals = coefs")


def test_noncanonical_falls_back_to_sweep(asg, oracle):
    """duplicates, out-of-range / even indices, non-finite coefficients: exact reference semantics."""
    OG = oracle.OracleGrid(2)
    OG.train(sinsum, 4, 4)
    # duplicate node + a node whose support lies partly outside [0,1] + an even index at l > 2
    OG.levels = np.vstack([OG.levels, OG.levels[5], [3, 3], [4, 3]])
    OG.indices = np.vstack([OG.indices, OG.indices[5], [5, 1], [2, 3]])
    OG.depths = OG.levels.sum(axis=1) - 1
    OG.coefs = np.concatenate([OG.coefs, [0.25, -0.5, 0.125]])
    OG.fvals = np.concatenate([OG.fvals, [0, 0, 0]])
    G = _mirror(asg, OG)
    rng = np.random.default_rng(0)
    X = np.vstack([rng.random((500, 2)) * 1.6 - 0.3, _edge_points(2, OG.lb, OG.ub, rng)])
    assert G.info()["canonical"] == 0 and G.info()["path"] == 1
    want, sc = OG.evaluate(X, return_abssum=True)
    assert_close(G.evaluate(X), want, sc, "non-canonical")
    with pytest.raises(asg.AsgError):
        G.set_path("subspace")
    # non-finite coefficient: Inf * 0 = NaN propagates exactly as in the reference (A8)
    OG2 = oracle.OracleGrid(2)
    OG2.train(sinsum, 3, 3)
    OG2.coefs[3] = np.inf
    G2 = _mirror(asg, OG2)
    X2 = rng.random((64, 2))
    w = OG2.evaluate(X2)
    g = G2.evaluate(X2)
    assert np.array_equal(np.isnan(g), np.isnan(w)) and np.any(np.isnan(w))
    assert G2.info()["path"] == 1
    # the same on a D = 3 grid, which has a BLOCK layout whose self-check must not see the non-finite sums; the
    # grid must also survive append (adapt!) and the packed round trip (round-1 advisor finding)
    for bad in (np.inf, np.nan):
        OG3 = oracle.OracleGrid(3)
        OG3.train(sinsum, 5, 5)
        OG3.coefs[3] = bad
        G3 = _mirror(asg, OG3)
        X3 = rng.random((256, 3))
        w3 = OG3.evaluate(X3)
        g3 = G3.evaluate(X3)
        assert G3.info()["path"] == 1
        assert np.array_equal(np.isnan(g3), np.isnan(w3)) and np.any(~np.isfinite(w3))
        assert np.array_equal(np.isinf(g3), np.isinf(w3)) and np.array_equal(np.sign(g3[np.isinf(g3)]), np.sign(w3[np.isinf(w3)]))
        ok = np.isfinite(w3)
        if ok.any():
            assert_close(g3[ok], w3[ok], None, "finite part next to a non-finite coefficient")
        H3 = asg.deserialize(asg.serialize(G3, packed=True))
        assert np.array_equal(H3.evaluate(X3), g3, equal_nan=True)


def test_invalid_nodes_rejected(asg):
    G = asg.SparseGridInterpolant(2)
    G.levels = np.array([[1, 1], [2, 1]], dtype=np.int64)
    G.indices = np.array([[1, 1], [1, 1]], dtype=np.int64)   # (2,1) is what phi throws on
    G.depths = np.array([1, 2])
    G.fvals = np.zeros(2)
    G.coefs = np.zeros(2)
    with pytest.raises(asg.InvalidNodeError):
        G.evaluate(np.zeros((1, 2)))
    G.levels = np.array([[1, 0]], dtype=np.int64)
    G.indices = np.array([[1, 1]], dtype=np.int64)
    G.depths, G.fvals, G.coefs = np.array([0]), np.zeros(1), np.zeros(1)
    with pytest.raises(asg.InvalidNodeError):
        G.sync()
    with pytest.raises(asg.InvalidNodeError):                # get_x rejects even index at l > 2
        asg.SparseGridInterpolant(1).nodes_x(np.array([[3]]), np.array([[2]]))


def test_empty_and_tiny(asg):
    G = asg.SparseGridInterpolant(3)
    assert G.evaluate(np.zeros((0, 3))).shape == (0,)
    X = np.random.default_rng(0).random((10, 3))
    assert np.all(G.evaluate(X) == 0.0)                      # empty grid: sum over nothing
    asg.rsg(G, 2, 2)
    G.coefs = np.arange(1.0, len(G) + 1)
    assert G.evaluate(np.full((1, 3), 0.5))[0] == 1.0        # only the root is non-zero at the centre
    assert asg.basis(G).shape == (7, 7)
    big = asg.SparseGridInterpolant(1)
    asg.train(big, lambda x: float(np.sin(3 * x[0])), 20, 20)   # 1-D, levels up to 20
    assert len(big) == 2 ** 19 + 1
    xs = np.random.default_rng(1).random((1000, 1))
    assert np.max(np.abs(big.evaluate(xs) - np.sin(3 * xs[:, 0]))) < 1e-10


def test_async_and_set_coefs(asg, cfg1):
    import ctypes as C
    G = _mirror(asg, cfg1)
    lib = asg._capi.load()
    X = np.random.default_rng(9).random((3000, 3))
    out = np.empty(3000)
    ev = C.c_void_p()
    asg._capi.check(lib.asg_eval_async(G._ensure(), C.c_void_p(X.ctypes.data), 3000, 3, 1, -1,
                                       C.c_void_p(out.ctypes.data), C.byref(ev)))
    asg._capi.check(lib.asg_wait(ev))
    assert np.array_equal(out, G.evaluate(X))
    new = np.linspace(-1, 1, len(G))
    asg._capi.check(lib.asg_grid_set_coefs(G._ensure(), 0, len(G), new.ctypes.data_as(asg._capi.f64p)))
    cfg = type(cfg1)(3)
    cfg.levels, cfg.indices, cfg.depths, cfg.coefs, cfg.fvals = cfg1.levels, cfg1.indices, cfg1.depths, new, cfg1.fvals
    want, sc = cfg.evaluate(X, return_abssum=True)
    assert_close(G.evaluate(X), want, sc, "after set_coefs")
    for p in ("sweep", "trie", "subspace"):
        G.set_path(p)
        assert_close(G.evaluate(X), want, sc, "after set_coefs " + p)


def test_full_size_properties(asg):
    """cfg3 at full grid size (D=10, 652 065 nodes): size-independent properties instead of the oracle.
    Linearity in the coefficients, interpolation at nodes of a trained analytic function, engines agree."""
    D = 10
    G = asg.SparseGridInterpolant(D)
    asg.rsg(G, 8, 8)
    assert len(G) == 652065
    info = G.info()
    assert info["subspaces"] == 19448 and info["canonical"] == 1
    rng = np.random.default_rng(2024)
    c1 = rng.standard_normal(len(G)) * 2.0 ** (-G.depths.astype(float))
    c2 = rng.standard_normal(len(G)) * 2.0 ** (-G.depths.astype(float))
    X = rng.random((20000, D))
    G.coefs = c1
    y1 = G.evaluate(X)
    G.coefs = c2
    y2 = G.evaluate(X)
    G.coefs = 2.0 * c1 - 0.5 * c2                            # exact scaling by powers of two
    y3 = G.evaluate(X)
    assert np.max(np.abs(y3 - (2.0 * y1 - 0.5 * y2))) < 1e-12
    # sweep engine and generic trie walk on a subsample agree with the record-free regular walk
    assert info["regular"] == 1
    G.set_path("sweep")
    ys = G.evaluate(X[:256])
    assert np.max(np.abs(ys - y3[:256])) < 1e-12
    G.set_path("trie")
    yt = G.evaluate(X[:4096])
    assert np.max(np.abs(yt - y3[:4096])) < 1e-12
    G.set_path("auto")
    # large batch (two points per thread) equals small batches (one point per thread)
    Xb = rng.random((400000, D))
    yb = G.evaluate(Xb)
    ysmall = np.concatenate([G.evaluate(Xb[k:k + 50000]) for k in range(0, 100000, 50000)])
    assert np.max(np.abs(yb[:100000] - ysmall)) < 1e-13
    # the same for the depth-masked walk and the fused surplus epilogue (train!'s call shape)
    fb = rng.standard_normal(len(Xb))
    for L in (3, 6):
        ym = G.evaluate(Xb, max_depth=L)
        ysm = G.evaluate(Xb[:50000], max_depth=L)
        assert np.max(np.abs(ym[:50000] - ysm)) < 1e-13
        assert np.max(np.abs(G.surplus(Xb, fb, max_depth=L) - (fb - ym))) < 1e-13
    G.set_path("sweep")
    assert np.max(np.abs(G.evaluate(Xb[:128], max_depth=6) - G.set_path("auto").evaluate(Xb, max_depth=6)[:128])) < 1e-12
    # coefficient 1 on the root only -> G == 1 everywhere inside; basis row sums to the constant
    G.coefs = np.where(G.depths == 1, 1.0, 0.0)
    assert np.all(G.evaluate(X[:1000]) == 1.0)


def test_hierarchize_forward_substitution(asg, oracle, cfg2):
    """coefs = E \\ fvals by depth-ordered forward substitution on the GPU."""
    G = asg.SparseGridInterpolant(3)
    asg.train(G, sinsum, 6, 6)
    want = G.coefs.copy()
    G.coefs = np.zeros(len(G))
    got = asg.hierarchize(G)
    assert_close(got, want, np.abs(G.fvals) + 1, "hierarchize == train coefficients")
    # against the explicit solve with the dehierarchization matrix (what the reference would do)
    import scipy.sparse.linalg as spla
    nodal = np.random.default_rng(0).standard_normal(len(G))
    E = asg.dehierarchization_matrix(G).tocsc()
    ref = spla.spsolve(E, nodal)
    got = asg.hierarchize(G, nodal).copy()
    assert_close(got, ref, 10 * (np.abs(nodal) + 1), "hierarchize == E \\ v")
    assert_close(G.evaluate(asg.stack(G)), nodal, np.abs(nodal) + 1, "interpolates the nodal values")
    # adaptive grid with hashed subspaces
    A = _mirror(asg, cfg2)
    A.coefs = np.zeros(len(A))
    assert_close(asg.hierarchize(A), cfg2.coefs, 4 * (np.abs(cfg2.fvals) + 1), "adaptive hierarchize")


@pytest.mark.parametrize("D", [11, 12, 13, 16])
def test_high_dimensions(asg, oracle, D):
    """D = 11, 12: subspace engine (largest instantiated kernels); D > 12: sweep engine fallback."""
    OG = oracle.OracleGrid(D)
    OG.rsg(3, 3)
    rng = np.random.default_rng(D)
    OG.coefs = rng.standard_normal(len(OG))
    G = _mirror(asg, OG)
    X = np.vstack([rng.random((3000, D)), _edge_points(D, OG.lb, OG.ub, rng, 16)])
    want, sc = OG.evaluate(X, return_abssum=True)
    assert_close(G.evaluate(X), want, sc, f"D={D}")
    assert G.info()["path"] == (2 if D <= 12 else 1)
    big = rng.random((400000, D)) if D <= 12 else None      # two points per thread path
    if big is not None:
        yb = G.evaluate(big)
        assert_close(yb[:2000], OG.evaluate(big[:2000]), None, f"D={D} large batch")
    B = asg.basis(X[:50], G).toarray()
    assert np.array_equal(B != 0, OG.basis(X[:50]) != 0)


def test_concurrent_handles_threads(asg, cfg1, cfg_nonunit):
    """different handles used from different host threads at the same time (per-handle mutex)"""
    import threading
    grids = [_mirror(asg, cfg1), _mirror(asg, cfg_nonunit), _mirror(asg, cfg1)]
    ogs = [cfg1, cfg_nonunit, cfg1]
    rng = np.random.default_rng(3)
    Xs = [og.lb + (og.ub - og.lb) * rng.random((20000, og.D)) for og in ogs]
    wants = [og.evaluate(X) for og, X in zip(ogs, Xs)]
    for G in grids:
        G._ensure()
    errs = []

    def work(i):
        try:
            for _ in range(5):
                got = grids[i].evaluate(Xs[i])
                if np.max(np.abs(got - wants[i])) > 1e-13:
                    errs.append(i)
        except Exception as e:  # pragma: no cover
            errs.append(repr(e))
    th = [threading.Thread(target=work, args=(i,)) for i in range(3)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs


def test_device_pointer_entry_points(asg, cfg1):
    """asg_eval_device / asg_surplus_device / asg_basis_dense_device with buffers resident in HBM (torch
    only supplies the device memory and the stream; no host copies inside the calls)."""
    import ctypes as C
    import torch
    G = _mirror(asg, cfg1)
    h = G._ensure()
    lib = asg._capi.load()
    rng = np.random.default_rng(17)
    Xh = rng.random((5000, 3))
    fh = rng.standard_normal(5000)
    want = cfg1.evaluate(Xh)
    X = torch.from_numpy(Xh).cuda()
    f = torch.from_numpy(fh).cuda()
    out = torch.empty(5000, dtype=torch.float64, device="cuda")
    side = torch.cuda.Stream()
    for stream in (torch.cuda.current_stream(), side):
        with torch.cuda.stream(stream):
            out.zero_()
            asg._capi.check(lib.asg_eval_device(h, C.c_void_p(X.data_ptr()), 5000, 3, 1, -1, C.c_void_p(out.data_ptr()),
                                                C.c_void_p(stream.cuda_stream)))
        stream.synchronize()
        assert_close(out.cpu().numpy(), want, None, "eval_device")
    ms = C.c_double()
    asg._capi.check(lib.asg_last_kernel_ms(C.byref(ms)))
    assert 0.0 < ms.value < 1000.0
    # column-major device matrix (Julia layout): strides (1, M)
    Xc = X.t().contiguous()
    asg._capi.check(lib.asg_eval_device(h, C.c_void_p(Xc.data_ptr()), 5000, 1, 5000, 4, C.c_void_p(out.data_ptr()), None))
    torch.cuda.synchronize()
    assert_close(out.cpu().numpy(), cfg1.evaluate(Xh, max_depth=4), None, "eval_device masked, column-major")
    asg._capi.check(lib.asg_surplus_device(h, C.c_void_p(X.data_ptr()), 5000, 3, 1, C.c_void_p(f.data_ptr()), -1,
                                           C.c_void_p(out.data_ptr()), None))
    torch.cuda.synchronize()
    assert_close(out.cpu().numpy(), fh - want, np.abs(fh) + 1, "surplus_device")
    N = len(G)
    Bd = torch.full((200, N), -7.0, dtype=torch.float64, device="cuda")
    asg._capi.check(lib.asg_basis_dense_device(h, C.c_void_p(X.data_ptr()), 200, 3, 1, 1e-16, 1,
                                               C.c_void_p(Bd.data_ptr()), N, 1, None))
    torch.cuda.synchronize()
    assert np.array_equal(Bd.cpu().numpy(), cfg1.basis(Xh[:200]))
    n0 = asg.launch_count()
    G.evaluate(Xh)
    assert asg.launch_count() > n0                       # the CUDA path is what ran


def test_fuzz_irregular_grids(asg, oracle):
    """random node subsets of regular grids in D = 1..7: mixtures of SUPER / per-group / generic records,
    dense and hashed subspaces, missing levels -- all three subspace walks and the sweep engine against the oracle"""
    rng = np.random.default_rng(2025)
    kinds = set()
    for trial in range(24):
        D = int(rng.integers(1, 8))
        OG = oracle.OracleGrid(D, lb=-rng.random(D), ub=1.0 + rng.random(D))
        OG.rsg(int(rng.integers(2, 7)), tuple(int(v) for v in rng.integers(2, 7, size=D)))
        keep = rng.random(len(OG)) < rng.choice([1.0, 0.95, 0.6, 0.25])
        keep[0] = True
        OG.levels, OG.indices, OG.depths = OG.levels[keep], OG.indices[keep], OG.depths[keep]
        OG.coefs = rng.standard_normal(len(OG.levels))
        OG.fvals = np.zeros(len(OG.levels))
        G = _mirror(asg, OG)
        info = G.info()
        assert info["canonical"] == 1
        kinds.add((info["regular"], info["hash_slots"] > 0))
        X = np.vstack([OG.lb + (OG.ub - OG.lb) * rng.random((1500, D)), _edge_points(D, OG.lb, OG.ub, rng, 8)])
        L = int(rng.integers(1, int(OG.depths.max()) + 1))
        want, sc = OG.evaluate(X, return_abssum=True)
        wantm, scm = OG.evaluate(X, max_depth=L, return_abssum=True)
        for path in ("subspace", "trie", "sweep"):
            G.set_path(path)
            assert_close(G.evaluate(X), want, sc, f"fuzz {trial} D={D} {path}")
            assert_close(G.evaluate(X, max_depth=L), wantm, scm, f"fuzz {trial} D={D} {path} depth<={L}")
        G.set_path("auto")
        B = asg.basis(X[:40], G).toarray()
        assert np.array_equal(B != 0, OG.basis(X[:40]) != 0)
        if trial % 4 == 0:  # large batch: the two-points-per-thread configuration on the same irregular grid
            Xb = OG.lb + (OG.ub - OG.lb) * rng.random((320000, D))
            yb = G.evaluate(Xb)
            wb, sb = OG.evaluate(Xb[:3000], return_abssum=True)
            assert_close(yb[:3000], wb, sb, f"fuzz {trial} D={D} large batch")
            assert_close(G.evaluate(Xb, max_depth=L)[:3000], OG.evaluate(Xb[:3000], max_depth=L), sb, "large batch masked")
    assert len(kinds) >= 3          # regular, irregular dense, irregular hashed were all exercised


def test_extrapolate_batched(asg, oracle, cfg1, cfg_nonunit):
    """batched linear extrapolation (1 + D batched evaluations) against the scalar oracle
    restatement of src/class.jl:695-737 (documented intent); affine functions are extrapolated exactly."""
    rng = np.random.default_rng(21)
    for OG in (cfg1, cfg_nonunit):
        G = _mirror(asg, OG)
        D = OG.D
        u = rng.random((400, D)) * 1.6 - 0.3                     # about half the coordinates outside [0,1]
        u[:50] = rng.random((50, D))                             # inside: plain evaluation
        u[50:60, 0] = 1.0                                        # on a face: not outside
        X = OG.lb + (OG.ub - OG.lb) * u
        got = asg.extrapolate(G, X)
        want = np.array([OG.extrapolate(x) for x in X])
        assert np.max(np.abs(got - want)) < 1e-11 * max(1.0, np.max(np.abs(want)))
        assert np.array_equal(got[:50], G.evaluate(X[:50]))
    aff = lambda x: float(1.0 + 2.0 * x[0] - x[1] + 0.5 * x[2])
    G = asg.SparseGridInterpolant(3)
    asg.train(G, aff, 4, 4)
    X = rng.random((200, 3)) * 3.0 - 1.0
    assert np.max(np.abs(asg.extrapolate(G, X) - (1.0 + 2.0 * X[:, 0] - X[:, 1] + 0.5 * X[:, 2]))) < 1e-12


def test_children_on_device(asg, oracle, cfg1, cfg2, cfg_nonunit):
    """adapt!'s candidate generation on the device (asg_children) against the host
    restatement of src/train.jl:337-398 -- the same rows, each once, in the reference's visiting order, with
    bit-exact coordinates; level-2 boundary nodes have a child on one side only (src/math.jl:931-959)."""
    from _asg_b200_pkg import interpolant as ip
    for OG in (cfg1, cfg2, cfg_nonunit):
        G = _mirror(asg, OG)
        for depth in sorted(set(int(d) for d in OG.depths)):
            parents = np.flatnonzero(OG.depths == depth)
            CLh, CIh = ip._candidates_host(OG.levels[parents], OG.indices[parents])
            key = np.concatenate([CLh, CIh], axis=1)
            _, first = np.unique(key, axis=0, return_index=True)
            first.sort()
            Lv, Ix, X, n_cand = G.children(depth)
            assert n_cand == 2 * OG.D * len(parents)
            assert np.array_equal(Lv, CLh[first]) and np.array_equal(Ix, CIh[first]), (OG.D, depth)
            assert np.array_equal(X, OG.stack(Lv, Ix))
    # both orders of work give the same grid
    f = lambda X: -1.0 / (0.1 + np.sum(X, axis=1))
    Ga, Gb = asg.SparseGridInterpolant(2), asg.SparseGridInterpolant(2)
    for G, dev in ((Ga, True), (Gb, False)):
        asg.train(G, f, 4, (3, 8), vectorized=True)
        asg.adapt(G, f, 12, tol=1e-4, verbose=False, vectorized=True, device_candidates=dev)
    assert np.array_equal(Ga.levels, Gb.levels) and np.array_equal(Ga.indices, Gb.indices)
    assert np.array_equal(Ga.coefs, Gb.coefs)


def test_packed_persistence(asg, oracle, cfg2, cfg_d10):
    """serialize(G, packed=True) carries the device-format blob; deserialize uploads it
    without packing again and evaluates bit-identically; corrupted / foreign blobs are rejected."""
    import ctypes as C
    rng = np.random.default_rng(8)
    for OG in (cfg2, cfg_d10):
        G = _mirror(asg, OG)
        X = OG.lb + (OG.ub - OG.lb) * rng.random((5000, OG.D))
        y = G.evaluate(X)
        dat = asg.serialize(G, packed=True)
        assert dat["levels"] is G.levels and dat["packed"].dtype == np.uint8 and dat["packed"].size > 16 * len(OG.levels)
        H = asg.deserialize(dat)
        assert H._dirty is False                               # no re-upload from the host arrays
        i0, i1 = G.info(), H.info()
        for k in ("N", "canonical", "subspaces", "coef_slots", "hash_slots", "block", "block_slots", "block_records", "regular"):
            assert i0[k] == i1[k], k
        for path in ("auto", "sweep"):
            G.set_path(path); H.set_path(path)
            assert np.array_equal(G.evaluate(X), H.evaluate(X))
        assert np.array_equal(y, H.set_path("auto").evaluate(X))
        G.set_path("auto")
        B0, B1 = asg.basis(X[:50], G), asg.basis(X[:50], H)
        assert (B0 != B1).nnz == 0
        # coefficient updates after an import reach every layout
        idx = np.arange(0, len(OG.levels), 7, dtype=np.int64)
        vals = rng.standard_normal(len(idx))
        for T in (G, H):
            T.coefs[idx] = vals
            asg._capi.check(asg._capi.load().asg_grid_scatter_coefs(T._ensure(), len(idx), idx.ctypes.data_as(asg._capi.i64p),
                                                                      vals.ctypes.data_as(asg._capi.f64p)))
        assert np.array_equal(G.evaluate(X), H.evaluate(X))
        # corrupted blob / wrong handle
        bad = dat["packed"].copy()
        bad[len(bad) // 2] ^= 0x40
        lib = asg._capi.load()
        assert lib.asg_grid_import(H._grid(), bad.ctypes.data_as(C.c_void_p), bad.size) == asg._capi.ASG_EINVAL
        assert "checksum" in lib.asg_last_error().decode()
        other = asg.SparseGridInterpolant(OG.D, lb=OG.lb - 1.0, ub=OG.ub)
        assert lib.asg_grid_import(other._grid(), dat["packed"].ctypes.data_as(C.c_void_p), dat["packed"].size) == asg._capi.ASG_EINVAL
        assert np.array_equal(G.evaluate(X), H.evaluate(X))    # the failed imports left H untouched


def test_basis_csc_on_device(asg, cfg1, cfg2, cfg_nonunit):
    """asg_basis_fill_csc: the CSR rows transposed on the device into the SparseMatrixCSC the reference returns
    (src/class.jl:483-503) -- identical entries, canonical form (rows ascending inside every column)."""
    rng = np.random.default_rng(17)
    for OG in (cfg1, cfg2, cfg_nonunit):
        G = _mirror(asg, OG)
        X = np.vstack([OG.lb + (OG.ub - OG.lb) * rng.random((700, OG.D)), OG.stack()[:300]])
        R = asg.basis(X, G, format="csr")
        Cm = asg.basis(X, G)
        assert Cm.format == "csc" and R.format == "csr" and Cm.shape == R.shape == (len(X), len(OG.levels))
        Rc = R.tocsc()
        Rc.sort_indices()
        assert np.array_equal(Cm.indptr, Rc.indptr) and np.array_equal(Cm.indices, Rc.indices)
        assert np.array_equal(Cm.data, Rc.data)
        assert np.array_equal(Cm.toarray(), OG.basis(X))
    E = asg.basis(_mirror(asg, cfg1))                      # basis(G): N x N, unit diagonal
    assert E.format == "csc" and np.all(E.diagonal() == 1.0)
    Z = asg.basis(np.full((3, 3), 7.5), _mirror(asg, cfg1))   # points outside the domain: empty matrix
    assert Z.nnz == 0 and Z.shape == (3, len(cfg1.levels))


def test_phi_exports(asg, oracle):
    """a4 and the scalar forms: phi(x), phi(x, l, i), phi(x, levels, indices) (src/math.jl:332-391) on the device,
    bit-identical to the oracle's restatement; invalid pairs raise like the reference (src/math.jl:355)."""
    rng = np.random.default_rng(12)
    xs = np.concatenate([rng.random(200) * 1.4 - 0.2, [0.0, 0.5, 1.0, -0.0, 0.25, 0.75, np.nan, np.inf, -np.inf, 1.0 + 1e-16]])
    t = np.concatenate([rng.random(100) * 3 - 1.5, [-1.0, 1.0, 0.0, np.nan]])
    assert np.array_equal(asg.phi(t), np.array([oracle.phi(v) for v in t]), equal_nan=True)
    assert asg.phi(0.25) == 0.75
    pairs = [(1, 1), (1, 7), (2, 0), (2, 2), (3, 1), (3, 3), (4, 5), (7, 33), (12, 1001), (5, 4), (30, 12345677)]
    for l, i in pairs:
        got = asg.phi(xs.reshape(-1, 1), np.full((len(xs), 1), l), np.full((len(xs), 1), i))
        want = np.array([oracle.phi(v, l, i) for v in xs])
        assert np.array_equal(got, want, equal_nan=True), (l, i)
        assert asg.phi(0.3, l, i) == oracle.phi(0.3, l, i)
    for l, i in [(0, 1), (2, 1), (-1, 0), (2, 3)]:
        with pytest.raises(ValueError):
            asg.phi(0.5, l, i)
        with pytest.raises(oracle.ArgumentError):
            oracle.phi(0.5, l, i)
    # vector form: one node, D dimensions, left fold; equals one entry of the basis matrix
    OG = oracle.OracleGrid(5)
    OG.rsg(5, 4)
    G = _mirror(asg, OG)
    X = rng.random((40, 5))
    B = OG.basis(X, dropzeros=False, atol=0.0)
    nodes = rng.choice(len(OG), size=40, replace=False)
    got = asg.phi(X, OG.levels[nodes], OG.indices[nodes])
    assert np.array_equal(got, B[np.arange(40), nodes])
    assert asg.phi(X[0], OG.levels[nodes[0]], OG.indices[nodes[0]], safe=True) == B[0, nodes[0]]
    with pytest.raises(AssertionError):
        asg.phi(X[0], OG.levels[0][:4], OG.indices[0][:4], safe=True)


def test_deserialize_ignores_stale_blob(asg, oracle, cfg1):
    """serialize() aliases the arrays (src/io.jl:30-38): a blob written before an in-place coefficient update, or a
    blob of another grid of the same shape, must not be installed (round-1 advisor finding)."""
    G = _mirror(asg, cfg1)
    dat = asg.serialize(G, packed=True)
    X = np.random.default_rng(3).random((500, 3))
    y0 = G.evaluate(X)
    dat["coefs"][:] = 2.0 * dat["coefs"]              # in-place update after the blob was written
    H = asg.deserialize(dat)
    assert np.array_equal(H.evaluate(X), 2.0 * y0)    # exact: scaling by a power of two
    # foreign blob: same D, domain and node count, other coefficients
    G2 = _mirror(asg, cfg1)
    G2.coefs = -cfg1.coefs
    dat2 = asg.serialize(G2)
    dat2["packed"] = asg.serialize(_mirror(asg, cfg1), packed=True)["packed"]
    assert np.array_equal(asg.deserialize(dat2).evaluate(X), -y0)   # exact: every term changes sign
    # a matching blob is still taken
    dat3 = asg.serialize(_mirror(asg, cfg1), packed=True)
    H3 = asg.deserialize(dat3)
    assert H3._dirty is False and np.array_equal(H3.evaluate(X), y0)


def test_extrapolate_empty_grid_asserts(asg):
    G = asg.SparseGridInterpolant(2)
    with pytest.raises(AssertionError):
        asg.extrapolate(G, np.array([[1.5, 0.5]]))


def test_host_pipeline_pieces_and_layouts():
    """The host-buffer pipeline (pieces, greedy windows, super-batches; csrc/asg_api.cu eval_super) with tiny piece /
    super-batch sizes so that a 30 000-point call runs through dozens of pieces and several super-batches: every
    host layout (row-major, Julia's column-major, D x n, generic strides), surplus, masked, against the oracle and
    bit-identical to each other."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r)
import asg_b200 as asg
from oracle import oracle as o
D = 5
OG = o.OracleGrid(D, lb=-np.arange(1, D + 1.0), ub=np.arange(1, D + 1.0))
OG.train(lambda x: float(np.sum(np.cos(x)) + x[0] * x[3]), 6, 5)
G = asg.SparseGridInterpolant(D, lb=OG.lb, ub=OG.ub)
G.levels, G.indices, G.depths, G.fvals, G.coefs = OG.levels, OG.indices, OG.depths, OG.fvals, OG.coefs
rng = np.random.default_rng(3)
M = 30011
X = OG.lb + (OG.ub - OG.lb) * rng.random((M, D))
X[::501, 1] = 99.0
want, sc = OG.evaluate(X, return_abssum=True)
a = G.evaluate(X)
err = np.abs(a - want)
assert np.all((err <= 1e-12 * np.abs(want)) | (err <= 1e-14 * np.maximum(1.0, sc)))
assert np.all(a[::501] == 0.0)
b = G.evaluate(np.asfortranarray(X))
c = G.evaluate(np.ascontiguousarray(X.T), dims=2)
d = G.evaluate(np.repeat(X, 2, axis=1)[:, ::2])
e = G.evaluate(np.repeat(X, 3, axis=0)[::3])
assert np.array_equal(a, b) and np.array_equal(a, c) and np.array_equal(a, d) and np.array_equal(a, e)
f = rng.standard_normal(M)
assert np.array_equal(G.surplus(X, f), f - a)
m = G.evaluate(X, max_depth=3)
wm = OG.evaluate(X, max_depth=3)
assert np.max(np.abs(m - wm)) < 1e-12
for path in ("sweep", "stream", "trie"):
    G.set_path(path)
    assert np.max(np.abs(G.evaluate(X) - a)) < 1e-12, path
print("pipeline ok")
''' % root
    env = dict(os.environ, ASG_PIPE_PIECE="1000", ASG_PIPE_SUPER="7777")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "pipeline ok" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_append_without_repacking(asg, oracle):
    """asg_grid_append (adapt!, src/train.jl:434-438): rows deeper than every node of the grid join the sweep arrays
    and are evaluated on top of the packed layouts; the packs happen when the appended part has grown large, before
    large batches and before basis / export -- not once per level. Same node set as the oracle's adapt!, same values."""
    OG = oracle.OracleGrid(2)
    OG.train(utility, 4, (3, 8))
    OG.adapt(utility, 20, tol=1e-5)
    G = asg.SparseGridInterpolant(2)
    asg.train(G, utility, 4, (3, 8))
    p0 = G.info()["packs"]
    log = []
    asg.adapt(G, utility, 20, tol=1e-5, verbose=False, log=log)
    info = G.info()
    levels_added = sum(1 for e in log if e.get("added"))
    assert levels_added >= 10
    assert info["packs"] - p0 <= 3 < levels_added           # round 1: one pack per level
    assert len(G) == len(OG) == 12119
    ko = np.lexsort(np.hstack([OG.levels, OG.indices]).T[::-1])
    kg = np.lexsort(np.hstack([G.levels, G.indices]).T[::-1])
    assert np.array_equal(OG.levels[ko], G.levels[kg]) and np.array_equal(OG.indices[ko], G.indices[kg])
    assert np.max(np.abs(OG.coefs[ko] - G.coefs[kg])) <= 1e-12
    # evaluation with appended nodes still outside the packed layouts: every engine, masked, surplus
    rng = np.random.default_rng(17)
    X = np.vstack([rng.random((5000, 2)), _edge_points(2, OG.lb, OG.ub, rng)])
    H = asg.SparseGridInterpolant(2)
    asg.train(H, utility, 4, (3, 8))
    asg.adapt(H, utility, 12, tol=1e-5, verbose=False)       # stops early: a few hundred appended nodes, no pack since
    assert H.info()["delta_nodes"] > 0
    OH = oracle.OracleGrid(2)
    OH.train(utility, 4, (3, 8))
    OH.adapt(utility, 12, tol=1e-5)
    assert len(H) == len(OH)
    k1 = np.lexsort(np.hstack([OH.levels, OH.indices]).T[::-1])
    k2 = np.lexsort(np.hstack([H.levels, H.indices]).T[::-1])
    OH.levels, OH.indices, OH.depths, OH.coefs, OH.fvals = H.levels, H.indices, H.depths, H.coefs, H.fvals  # same order as H
    want, sc = OH.evaluate(X, return_abssum=True)
    for path in ("auto", "trie", "stream", "sweep"):
        H.set_path(path)
        assert_close(H.evaluate(X), want, sc, f"appended nodes, {path}")
        assert_close(H.evaluate(X, max_depth=9), OH.evaluate(X, max_depth=9), sc, f"appended nodes, masked, {path}")
    H.set_path("auto")
    f = rng.standard_normal(len(X))
    assert_close(H.surplus(X, f), f - want, sc + np.abs(f), "appended nodes, surplus")
    assert H.info()["delta_nodes"] > 0                       # none of this packed
    # coefficient updates reach appended nodes
    nd = H.info()["delta_nodes"]
    idx = np.ascontiguousarray(np.arange(len(H) - nd, len(H), dtype=np.int64)[::3])
    vals = np.ascontiguousarray(rng.standard_normal(len(idx)))
    lib = asg._capi.load()
    asg._capi.check(lib.asg_grid_scatter_coefs(H._ensure(), len(idx), idx.ctypes.data_as(asg._capi.i64p),
                                               vals.ctypes.data_as(asg._capi.f64p)))
    H.coefs[idx] = vals
    OH.coefs = H.coefs
    want2, sc2 = OH.evaluate(X, return_abssum=True)
    assert_close(H.evaluate(X), want2, sc2, "appended nodes after scatter_coefs")
    # basis folds them in first; the matrix then has a column for every node
    B = asg.basis(X[:200], H)
    assert H.info()["delta_nodes"] == 0 and B.shape == (200, len(H))
    assert_close(B @ H.coefs, want2[:200], sc2[:200], "basis after the fold")
    assert_close(H.evaluate(X), want2, sc2, "after the fold")
    # rows that are NOT deeper than the grid (possible repeats) take the full pack at once, like before
    p1 = H.info()["packs"]
    asg._capi.check(lib.asg_grid_append(H._ensure(), 1, np.array([[2, 1]], dtype=np.int64).ctypes.data_as(asg._capi.i64p),
                                        np.array([[0, 1]], dtype=np.int64).ctypes.data_as(asg._capi.i64p), 2, 1,
                                        np.array([0.5]).ctypes.data_as(asg._capi.f64p)))
    assert H.info()["packs"] == p1 + 1 and H.info()["canonical"] == 0      # a repeated node: the sweep engine takes over


def test_fuzz_engines_against_oracle(asg, oracle):
    """benchmarks/fuzz_paths.py compares the engines with each other (good at layout bugs, blind to a shared misreading
    of the reference); this is the same kind of random grid -- dimension, depth, level caps, node subsets, domains,
    adapt!-grown -- with every engine compared to the ORACLE's long-double sum, unmasked and depth-masked."""
    rng = np.random.default_rng(20261020)
    done = 0
    for t in range(60):
        D = int(rng.integers(1, 11))
        depth = int(rng.integers(2, {1: 11, 2: 10, 3: 9, 4: 8}.get(D, 6 if D <= 7 else 5)))
        caps = tuple(int(v) for v in rng.integers(2, depth + 2, size=D)) if rng.random() < 0.5 else depth
        lb = np.zeros(D) if rng.random() < 0.4 else -rng.random(D) * 3
        ub = np.ones(D) if rng.random() < 0.4 else lb + 0.5 + rng.random(D) * 4
        OG = oracle.OracleGrid(D, lb=lb, ub=ub)
        OG.rsg(depth, caps)
        if len(OG) > 60000:
            continue
        mode = rng.choice(["full", "subset", "adapt"])
        if mode == "subset":
            keep = rng.random(len(OG)) < rng.choice([0.97, 0.8, 0.5])
            keep[0] = True
            OG.levels, OG.indices, OG.depths = OG.levels[keep], OG.indices[keep], OG.depths[keep]
        OG.coefs = rng.standard_normal(len(OG.levels)) * 2.0 ** (-rng.random() * OG.depths)
        OG.fvals = np.zeros(len(OG.levels))
        G = _mirror(asg, OG)
        if mode == "adapt" and D <= 4 and len(OG) < 3000:
            c = rng.random(D)
            f = lambda x, lb=lb, ub=ub, c=c: float(np.exp(-np.sum(((np.asarray(x) - lb) / (ub - lb) - c) ** 2) * 8.0))
            tol = 10.0 ** -int(rng.integers(2, 4))
            extra = int(rng.integers(1, 3))
            OG.train(f, depth, caps)
            OG.adapt(f, depth + extra, tol=tol)
            G = asg.SparseGridInterpolant(D, lb=lb, ub=ub)
            asg.train(G, f, depth, caps)
            asg.adapt(G, f, depth + extra, tol=tol, verbose=False)
            assert len(G) == len(OG), (t, D, depth)
            OG.levels, OG.indices, OG.depths, OG.coefs, OG.fvals = G.levels, G.indices, G.depths, G.coefs, G.fvals
        M = 400
        u = rng.random((M, D))
        u[rng.random((M, D)) < 0.05] = rng.integers(0, 65) / 64.0          # knots
        u[::37, rng.integers(0, D)] = rng.choice([1.0, 0.0, 1.25, -0.5, np.nan])
        X = lb + (ub - lb) * u
        L = int(rng.integers(1, int(OG.depths.max()) + 1))
        want, sc = OG.evaluate(X, return_abssum=True)
        wantm = OG.evaluate(X, max_depth=L)
        info = G.info()
        paths = ["sweep"] + (["stream", "trie"] + (["block"] if info["block"] else []) if info["canonical"] else [])
        for p in paths:
            G.set_path(p)
            assert_close(G.evaluate(X), want, sc, f"grid {t} D={D} depth={depth} {mode} {p}")
            assert_close(G.evaluate(X, max_depth=L), wantm, sc, f"grid {t} D={D} depth={depth} {mode} {p} depth<={L}")
        done += 1
    assert done >= 40


def test_pageable_staging_boundaries(asg, oracle):
    """Pageable host buffers go through the library's pinned staging from 64 MiB of points / 8 MiB of results on
    (asg_api.cu, eval_super / d2h_fast): sizes just below and above both limits, row- and column-major, evaluate and
    surplus, against the same points evaluated in small calls (which take the direct path) and against the oracle."""
    G = asg.SparseGridInterpolant(4)
    asg.train(G, lambda X: np.sum(np.sin(2 * np.pi * X), axis=1), 7, 6, vectorized=True)
    OG = oracle.OracleGrid(4)
    OG.levels, OG.indices, OG.depths, OG.coefs, OG.fvals = G.levels, G.indices, G.depths, G.coefs, G.fvals
    rng = np.random.default_rng(11)
    Mmax = 3_000_001
    X = rng.random((Mmax, 4))
    X[::1013, 2] = 1.0
    X[::7919, 0] = -0.5
    f = rng.standard_normal(Mmax)
    ref = np.concatenate([G.evaluate(X[i:i + 200_000]) for i in range(0, Mmax, 200_000)])
    sel = rng.choice(Mmax, size=500, replace=False)
    want, sc = OG.evaluate(X[sel], return_abssum=True)
    assert_close(ref[sel], want, sc, "small calls vs oracle")
    # 64 MiB of D = 4 points are 2 097 152 points; 8 MiB of results 1 048 576 points
    for M in (1_048_575, 1_048_577, 2_097_151, 2_097_153, Mmax):
        Xm = np.ascontiguousarray(X[:M])
        y = G.evaluate(Xm)
        assert np.array_equal(y, ref[:M]), f"row-major, M = {M}"
        yf = G.evaluate(np.asfortranarray(Xm))
        assert np.array_equal(yf, ref[:M]), f"column-major, M = {M}"
        a = G.surplus(Xm, f[:M], max_depth=5)
        assert np.array_equal(a, f[:M] - G.evaluate(Xm, max_depth=5)), f"surplus, M = {M}"
