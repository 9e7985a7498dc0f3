"""BLOCK layout (csrc/asg_block.h): packer checks on the CPU, parity of the compiled walk on the GPU.

The BLOCK walk is the default evaluation kernel for grids that are (nearly) regular in their last three
dimensions -- BASELINE.json configs[2] (D = 10, 652 065 nodes) among them; the flat group stream stays the
kernel of every other canonical grid, so both are run against the oracle here on identical inputs.
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import assert_close, sinsum


def _pack_info(asg, levels, indices):
    gi = asg._capi.GridInfo()
    Lv = np.ascontiguousarray(levels, dtype=np.int64)
    Ix = np.ascontiguousarray(indices, dtype=np.int64)
    rc = asg._capi.load().asg_pack_info(Lv.shape[1], len(Lv), Lv.ctypes.data_as(asg._capi.i64p),
                                        Ix.ctypes.data_as(asg._capi.i64p), Lv.shape[1], 1, C.byref(gi))
    return rc, gi.as_dict(), asg._capi.load().asg_last_error().decode()


def test_block_layout_regular_grids(asg):
    """asg_pack_info runs the packer's self-check (host model of the kernel's walk against the direct sum
    over the nodes, with and without a depth mask) for every grid of at most 50 000 nodes, and on every pack follows the
    DEVICE form of the record stream (tile headers, block_device_stream_ok) from tile to tile the way the kernel does."""
    for D, depth, ml, nrec in [(4, 6, 6, 1), (5, 5, 5, 1), (6, 7, (6, 7, 8, 9, 10, 11), None), (10, 5, 5, None),
                               (7, 4, (2, 3, 4, 4, 2, 3, 4), None), (12, 3, 3, None)]:
        G = asg.SparseGridInterpolant(D)
        asg.rsg(G, depth, ml)
        rc, info, msg = _pack_info(asg, G.levels, G.indices)
        assert rc == 0, msg
        assert info["block"] == 1, msg
        if nrec is not None:
            assert info["block_slots"] == len(G)          # no padding unless level caps cut the last three dimensions
            assert info["block_records"] == nrec          # the whole grid is one compiled 4-D / 5-D suffix
        else:
            assert len(G) <= info["block_slots"] <= 2 * len(G) + 4096
    # the benchmark grid: C(14, 7) = 3432 prefix vectors in 619 records (small budgets take 4..7-D suffixes), no padding
    os.environ["ASG_PACK_SELFCHECK"] = "1"
    G = asg.SparseGridInterpolant(10)
    asg.rsg(G, 8, 8)
    rc, info, msg = _pack_info(asg, G.levels, G.indices)
    assert rc == 0 and info["block"] == 1 and info["block_slots"] == 652065 and info["block_records"] == 619, msg
    assert info["block_stack"] == 5 and info["block_tiles"] == 182
    # D = 3: no prefix at all, one record
    G = asg.SparseGridInterpolant(3)
    asg.rsg(G, 7, 5)
    rc, info, msg = _pack_info(asg, G.levels, G.indices)
    assert rc == 0 and info["block"] == 1 and info["block_records"] == 1 and info["block_slots"] == 1073, msg  # caps: padding
    # BASELINE configs[4] (D = 4, depth 10): suffix budgets up to 9 since round 2 (the WIDE kernels), never paired
    G = asg.SparseGridInterpolant(4)
    asg.rsg(G, 10, 10)
    rc, info, msg = _pack_info(asg, G.levels, G.indices)
    assert rc == 0 and info["block"] == 1 and info["block_slots"] == 46721 and info["block_pair"] == 0, msg
    # outside the layout's limits: D < 3, suffix levels > 10
    for D, depth in [(2, 5), (4, 12)]:
        G = asg.SparseGridInterpolant(D)
        asg.rsg(G, depth, depth)
        rc, info, msg = _pack_info(asg, G.levels, G.indices)
        assert rc == 0 and info["canonical"] == 1 and info["block"] == 0 and "BLOCK" in msg


def test_block_layout_irregular_grids(asg, oracle):
    """node subsets (padding with zeros, virtual parent records) and an adapt! grid"""
    rng = np.random.default_rng(77)
    took = 0
    for trial in range(20):
        D = int(rng.integers(4, 9))
        G = asg.SparseGridInterpolant(D)
        asg.rsg(G, int(rng.integers(3, 7)), tuple(int(v) for v in rng.integers(2, 7, size=D)))
        keep = rng.random(len(G)) < rng.choice([0.97, 0.8, 0.55])
        keep[0] = True
        rc, info, msg = _pack_info(asg, G.levels[keep], G.indices[keep])
        assert rc == 0, msg                               # a failed self-check would be ASG_ESTATE
        assert info["canonical"] == 1
        if info["block"]:
            took += 1
            assert keep.sum() <= info["block_slots"] <= 2 * keep.sum() + 4096
    assert took >= 5
    OG = oracle.OracleGrid(4)
    OG.train(sinsum, 5, 5)
    OG.adapt(sinsum, 8, tol=3e-2)
    rc, info, msg = _pack_info(asg, OG.levels, OG.indices)
    assert rc == 0 and info["canonical"] == 1, msg


# ---------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------
def _mirror(asg, OG):
    G = asg.SparseGridInterpolant(OG.D, lb=OG.lb, ub=OG.ub)
    G.levels, G.indices, G.depths = OG.levels.copy(), OG.indices.copy(), OG.depths.copy()
    G.fvals, G.coefs = OG.fvals.copy(), OG.coefs.copy()
    return G


def _edge_points(D, lb, ub, rng):
    from test_gpu_parity import _edge_points as e
    return e(D, lb, ub, rng, 32)


def _grids(oracle):
    rng = np.random.default_rng(99)
    out = []
    for D, depth, ml, keepfrac, unit in [(3, 8, 8, 1.0, False), (3, 7, (5, 6, 4), 0.8, True), (4, 7, 7, 1.0, True), (10, 5, 5, 1.0, True), (6, 6, (4, 5, 6, 3, 6, 6), 1.0, False),
                                         (5, 6, 6, 0.9, False), (7, 4, 4, 0.7, True), (4, 8, 8, 1.0, False),
                                         (12, 3, 3, 1.0, True),
                                         (3, 10, 10, 1.0, True), (4, 10, 10, 1.0, False)]:   # suffix budgets 8 and 9 (levels up to 10)
        lb = np.zeros(D) if unit else -rng.random(D)
        ub = np.ones(D) if unit else 1.0 + rng.random(D)
        OG = oracle.OracleGrid(D, lb=lb, ub=ub)
        OG.rsg(depth, ml)
        if keepfrac < 1.0:
            keep = rng.random(len(OG)) < keepfrac
            keep[0] = True
            OG.levels, OG.indices, OG.depths = OG.levels[keep], OG.indices[keep], OG.depths[keep]
        OG.coefs = rng.standard_normal(len(OG.levels)) * 2.0 ** (-OG.depths.astype(float) / 2)
        OG.fvals = np.zeros(len(OG.levels))
        out.append(OG)
    # 3-D chunks larger than a shared-memory tile: split into S2 records (level of dimension D-3 as the stage)
    OG = oracle.OracleGrid(4)
    OG.rsg(11, 11)
    keep = (np.sum(OG.levels[:, 1:] - 1, axis=1) <= 7) & (OG.levels[:, 0] <= 5)
    OG.levels, OG.indices, OG.depths = OG.levels[keep], OG.indices[keep], OG.depths[keep]
    OG.coefs = rng.standard_normal(len(OG.levels))
    OG.fvals = np.zeros(len(OG.levels))
    out.append(OG)
    return out


@pytest.mark.gpu
def test_block_walk_parity(asg, oracle):
    rng = np.random.default_rng(5)
    for OG in _grids(oracle):
        D = OG.D
        G = _mirror(asg, OG)
        info = G.info()
        assert info["block"] == 1, (D, len(OG.levels))
        if D == 4 and int(OG.levels[:, 0].max()) == 5 and int(OG.depths.max()) >= 10:
            assert info["block_records"] > 5              # 5 prefix levels, the fine ones split along dimension D-3
        X = np.vstack([OG.lb + (OG.ub - OG.lb) * rng.random((3000, D)), _edge_points(D, OG.lb, OG.ub, rng)])
        want, sc = OG.evaluate(X, return_abssum=True)
        L = int(rng.integers(1, int(OG.depths.max()) + 1))
        wantm, scm = OG.evaluate(X, max_depth=L, return_abssum=True)
        res = {}
        for path in ("block", "stream"):
            G.set_path(path)
            n0 = asg.launch_count()
            got = G.evaluate(X)
            assert asg.launch_count() > n0
            assert_close(got, want, sc, f"D={D} {path}")
            with np.errstate(invalid="ignore"):
                u = (X - OG.lb) / (OG.ub - OG.lb)             # the reference tests the SCALED coordinate (math.jl:308, 349-353)
                outside = ~np.all((u >= 0.0) & (u <= 1.0), axis=1)
            assert outside.sum() >= 20 and np.all(got[outside] == 0.0)
            gm = G.evaluate(X, max_depth=L)
            assert_close(gm, wantm, scm, f"D={D} {path} depth<={L}")
            res[path] = got
        # surplus = f - G through the same kernel; coefficient updates must reach the BLOCK copy too
        G.set_path("block")
        f = rng.standard_normal(len(X))
        assert_close(G.surplus(X, f), f - want, sc + np.abs(f), f"D={D} surplus")
        idx = rng.choice(len(OG.levels), size=min(50, len(OG.levels)), replace=False)
        OG.coefs[idx] = rng.standard_normal(len(idx))
        G.coefs[idx] = OG.coefs[idx]
        lib = asg._capi.load()
        idx64 = np.ascontiguousarray(idx, dtype=np.int64)
        vals = np.ascontiguousarray(OG.coefs[idx])
        asg._capi.check(lib.asg_grid_scatter_coefs(G._ensure(), len(idx64), idx64.ctypes.data_as(asg._capi.i64p),
                                                   vals.ctypes.data_as(asg._capi.f64p)))
        want2, sc2 = OG.evaluate(X, return_abssum=True)
        assert_close(G.evaluate(X), want2, sc2, f"D={D} block after scatter_coefs")
        G.set_path("stream")
        assert_close(G.evaluate(X), want2, sc2, f"D={D} stream after scatter_coefs")


@pytest.mark.gpu
def test_block_walk_large_sorted_batch(asg, oracle):
    """>= 2^20 points: two points per thread and the cell-key ordering of the points (asg_sort.cu). The
    per-point arithmetic does not depend on the batch, so the results equal those of a small batch bit for bit."""
    rng = np.random.default_rng(11)
    OG = [g for g in _grids(oracle) if g.D == 10][0]         # D = 10, depth 5
    G = _mirror(asg, OG)
    G.set_path("block")
    M = (1 << 20) + 12345
    X = rng.random((M, 10))
    X[::1000, 3] = 1.5                                       # some out-of-domain points
    X[7::5000, 9] = np.nan
    y = G.evaluate(X)
    sel = np.concatenate([np.arange(0, 3000), np.arange(M - 3000, M)])
    want, sc = OG.evaluate(X[sel], return_abssum=True)
    assert_close(y[sel], want, sc, "large sorted batch")
    assert np.array_equal(y[sel], G.evaluate(X[sel]))
    assert np.all(y[::1000] == 0.0) and np.all(y[7::5000] == 0.0)
    ym = G.evaluate(X, max_depth=3)
    assert_close(ym[sel], OG.evaluate(X[sel], max_depth=3), sc, "large sorted batch, masked")
    # column-major points (Julia's M x D matrices) take the same path
    yf = G.evaluate(np.asfortranarray(X))
    assert np.array_equal(y, yf, equal_nan=True)
    G.set_path("stream")
    ys = G.evaluate(X[sel])
    assert_close(ys, want, sc, "stream")


@pytest.mark.gpu
def test_block_walk_train_cfg4(asg, oracle):
    """train! on the D = 6 anisotropic grid of BASELINE.json configs[3] (15 089 nodes): every depth's surplus
    step runs the masked BLOCK walk; coefficients must match the oracle's"""
    OG = oracle.OracleGrid(6)
    OG.train(sinsum, 7, tuple(range(6, 12)))
    G = asg.SparseGridInterpolant(6)
    asg.train(G, sinsum, 7, tuple(range(6, 12)))
    assert G.info()["block"] == 1
    assert np.array_equal(G.levels, OG.levels) and np.array_equal(G.indices, OG.indices)
    assert np.max(np.abs(G.coefs - OG.coefs)) < 1e-13


@pytest.mark.gpu
def test_block_walk_adapt_cfg4(asg):
    """adapt! on BASELINE configs[3] (D = 6, depth 13, tol 3e-3): the grids it grows are irregular (zero padding,
    stage-only parent records, split records); every level's candidate surpluses come from the BLOCK walk while the
    layout applies. The node count is the one SURVEY.md section 8 derives (19 929, converged at depth 12) and the
    three engines agree on the final grid."""
    def f(X):
        return np.sum(np.sin(2 * np.pi * X), axis=1)
    G = asg.SparseGridInterpolant(6)
    asg.train(G, f, 7, tuple(range(6, 12)), vectorized=True)
    assert len(G) == 15089 and G.info()["block"] == 1
    log = []
    asg.adapt(G, f, 13, tol=3e-3, verbose=False, vectorized=True, log=log)
    assert len(G) == 19929, log
    rng = np.random.default_rng(4)
    X = np.vstack([rng.random((20000, 6)), rng.integers(0, 4097, size=(20000, 6)) / 4096.0])
    G.set_path("sweep")
    want = G.evaluate(X)
    for path in ("stream", "block"):
        try:
            G.set_path(path)
        except asg.AsgError:
            assert path == "block"       # the final grid may need more padding than the layout accepts
            continue
        assert np.max(np.abs(G.evaluate(X) - want)) < 5e-14, path


@pytest.mark.gpu
def test_engines_differential_fuzz():
    """benchmarks/fuzz_paths.py: 80 random grids (D = 1..12, caps, node subsets, non-unit domains, adapt!-grown), every
    engine against the sweep engine, unmasked and depth-masked, batches of 257 .. 2^20 points"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "benchmarks", "fuzz_paths.py"), "80", "12345"], cwd=root,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "all grids agree" in r.stdout


@pytest.mark.gpu
def test_block_walk_pair_mode(asg, oracle):
    """PAIR mode (asg_kernels_block.cuh): two neighbours of the cell-key order per thread, one coefficient load for both
    wherever every level on the path is <= 4. Regular and adapt!-grown grids, D = 3 .. 10, batches where nearly every
    pair shares its cell (D small) and where nearly none does (D = 10: the follow-up launch does the work); unmasked,
    masked and surplus; the results equal the unpaired walk's bit for bit and the oracle's within tolerance."""
    rng = np.random.default_rng(21)
    paired = 0
    cfg4 = oracle.OracleGrid(6)
    cfg4.rsg(7, tuple(range(6, 12)))             # 15 089 nodes, 923 subspaces, suffix levels up to 8
    cfg4.coefs = rng.standard_normal(len(cfg4)) * 2.0 ** (-cfg4.depths.astype(float))
    for OG in _grids(oracle) + [cfg4]:
        D = OG.D
        G = _mirror(asg, OG)
        if G.info()["block"] != 1:
            continue
        G.set_path("block")
        M = 300001
        X = OG.lb + (OG.ub - OG.lb) * rng.random((M, D))
        X[::777] = OG.lb + (OG.ub - OG.lb) * (rng.integers(0, 65, size=(len(X[::777]), D)) / 64.0)   # knots
        X[5::1001, 0] = OG.ub[0] + 1.0
        X[9::2003, D - 1] = np.nan
        f = rng.standard_normal(M)
        L = int(rng.integers(2, int(OG.depths.max()) + 1))
        os.environ["ASG_BLOCK_PAIR"] = "0"
        try:
            y0 = G.evaluate(X)
            m0 = G.evaluate(X, max_depth=L)
            s0 = G.surplus(X, f)
        finally:
            del os.environ["ASG_BLOCK_PAIR"]
        os.environ["ASG_BLOCK_PAIR"] = "1e-9"
        try:
            n0 = asg.launch_count()
            y1 = G.evaluate(X)
            if G.info()["subspaces"] >= 512 and G.info()["block_pair"] == 1:   # (smaller grids skip the point ordering)
                assert asg.launch_count() - n0 >= 3      # keys + paired walk + follow-up walk
                paired += 1
            m1 = G.evaluate(X, max_depth=L)
            s1 = G.surplus(X, f)
        finally:
            del os.environ["ASG_BLOCK_PAIR"]
        assert np.array_equal(y0, y1, equal_nan=True), D
        assert np.array_equal(m0, m1, equal_nan=True), (D, L)
        assert np.array_equal(s0, s1, equal_nan=True), D
        sel = rng.choice(M, size=2000, replace=False)
        want, sc = OG.evaluate(X[sel], return_abssum=True)
        assert_close(y1[sel], want, sc, f"pair mode D={D}")
        assert np.all(y1[5::1001] == 0.0) and np.all(y1[9::2003] == 0.0)
    assert paired >= 2   # D = 10 depth 5 and the D = 6 anisotropic grid of BASELINE configs[3]
