"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, fails
loudly without a GPU, and the host logic kept on the CPU (node enumeration, children, acceptance,
packing) matches the oracle. No compute call is made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "asg_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(asg_[a-z0-9_]+)\s*\(", src)))


def test_abi_exports_every_declared_symbol(asg):
    names = _declared_symbols()
    assert len(names) >= 30 and "asg_eval" in names and "asg_basis_fill" in names
    lib = C.CDLL(asg.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/asg_cuda.h but not exported"
    bound = set(asg._capi.SIGNATURES) | set(asg._capi.STRING_GETTERS)
    assert bound == set(names), f"ctypes table out of step with the header: {bound ^ set(names)}"
    assert b"sm_100a" in asg._capi.load().asg_version()


def test_no_cpu_fallback(asg):
    """Without a device every compute entry point fails with ASG_ENODEVICE (the product has no CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(asg.AsgError) as e:
        asg.init()
    assert e.value.code == asg._capi.ASG_ENODEVICE
    G = asg.SparseGridInterpolant(2)
    asg.rsg(G, 3, 3)
    with pytest.raises(asg.AsgError):
        G.evaluate(np.zeros((4, 2)))
    with pytest.raises(asg.AsgError):
        asg.stack(G)
    lib = asg._capi.load()
    out = np.zeros(1)
    assert lib.asg_eval(None, None, 0, 1, 1, -1, None) == asg._capi.ASG_EINVAL
    assert lib.asg_measure_fp64_peak(out.ctypes.data_as(asg._capi.f64p)) == asg._capi.ASG_ENODEVICE
    assert b"no CPU fallback" in lib.asg_last_error() or b"asg_init" in lib.asg_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "adaptivesg.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".jl")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in txt.lower() or f == "README.md", f"{f} mentions the oracle"
    assert "oracle" not in open(os.path.join(ROOT, "asg_b200.py")).read()


def test_constructor_asserts(asg):
    with pytest.raises(AssertionError):
        asg.SparseGridInterpolant(0)
    with pytest.raises(AssertionError):
        asg.SparseGridInterpolant(2, lb=[0, 0], ub=[1, 0])
    with pytest.raises(AssertionError):
        asg.SparseGridInterpolant(2, lb=[0, 0, 0])
    with pytest.raises(AssertionError):
        asg.SparseGridInterpolant(1, lb=[-np.inf])
    G = asg.SparseGridInterpolant(3, lb=[1, 2, 3], ub=[2, 3, 4])
    assert len(G) == 0 and G.ndims == 3 and G.size == (0, 3) and G.levels.shape == (0, 3)
    assert [1.5, 2.5, 3.5] in G and [0.5, 2.5, 3.5] not in G
    assert G.clamp([0, 9, 3.5]).tolist() == [1.0, 3.0, 3.5]
    assert G.rand(5, np.random.default_rng(0)).shape == (3, 5)          # D x n like the reference
    assert "R^3" in repr(G) and asg.maxlevels(G).tolist() == [0, 0, 0]
    lib = asg._capi.load()
    h = C.c_void_p()
    bad = np.array([0.0, np.nan])
    assert lib.asg_grid_create(2, bad.ctypes.data_as(asg._capi.f64p), None, C.byref(h)) == asg._capi.ASG_EINVAL
    assert lib.asg_grid_create(0, None, None, C.byref(h)) == asg._capi.ASG_EINVAL
    assert lib.asg_grid_create(65, None, None, C.byref(h)) == asg._capi.ASG_EUNSUPPORTED
    assert lib.asg_grid_create(2, None, None, C.byref(h)) == 0 and h.value
    assert lib.asg_grid_destroy(h) == 0


def test_rsg_matches_oracle(asg, oracle):
    for D, md, ml in [(3, 7, 5), (2, 4, (3, 8)), (6, 5, tuple(range(6, 12))), (1, 6, 9), (4, 3, 1), (5, 9, (2, 3, 4, 3, 2))]:
        G = asg.SparseGridInterpolant(D)
        asg.rsg(G, md, ml)
        O = oracle.OracleGrid(D)
        O.rsg(md, ml)
        assert np.array_equal(G.levels, O.levels) and np.array_equal(G.indices, O.indices)
        assert np.array_equal(G.depths, O.depths) and not G.coefs.any() and not G.fvals.any()
    with pytest.raises(ValueError):
        asg.rsg(G, 1, 3)
    with pytest.raises(ValueError):
        asg.rsg(G, 3, 0)


def test_children_validity_acceptance_match_oracle(asg, oracle):
    rng = np.random.default_rng(1)
    I = asg.interpolant
    for l, i in [(1, 1), (2, 0), (2, 2), (3, 1), (3, 3), (7, 45), (12, 1023)]:
        for side in ("left", "right"):
            nl, ni = I._children(np.array([l]), np.array([i]), side == "right")
            assert (int(nl[0]), int(ni[0])) == oracle.child1(l, i, side)
            L, Ix = asg.child([1, l, 1], [1, i, 1], 2, side)
            assert (int(L[1]), int(Ix[1])) == oracle.child1(l, i, side)
    with pytest.raises(ValueError):
        I._children(np.array([2]), np.array([1]), False)
    for _ in range(300):
        L = rng.integers(0, 5, size=3)
        Ix = rng.integers(-1, 9, size=3)
        assert asg.isvalid(L, Ix) == oracle.isvalid(L, Ix)
    assert not asg.isvalid([1, 2], [1])
    a = rng.standard_normal(200) * 1e-3
    f = np.where(rng.random(200) < 0.3, rng.standard_normal(200) * 1e-4, rng.standard_normal(200))
    for use_r in (False, True):
        got = asg.isbig2add(a, f, 1e-3, 1e-3, use_r)
        want = [oracle.isbig2add(x, y, 1e-3, 1e-3, use_r) for x, y in zip(a, f)]
        assert got.tolist() == want
    assert asg.nodedepth([3, 1, 2]) == 4 and asg.allindices_of_level(4).tolist() == [1, 3, 5, 7]
    with pytest.raises(ValueError):
        asg.allindices_of_level(0)


def test_serialize_roundtrip_aliases(asg):
    G = asg.SparseGridInterpolant(2, lb=[0, 1], ub=[2, 3])
    asg.rsg(G, 3, 3)
    dat = asg.serialize(G)
    assert set(dat) == {"D", "lb", "ub", "levels", "indices", "depths", "fvals", "coefs"}
    assert dat["levels"] is G.levels and dat["coefs"] is G.coefs          # aliased, not copied (src/io.jl:33-38)
    G2 = asg.deserialize(dat)
    assert G2.lb == G.lb and G2.ub == G.ub and G2.levels is G.levels and len(G2) == len(G)
    assert asg.coefficients(G2) is G2.coefs
    with pytest.raises(AssertionError):
        asg.deserialize({"D": 2, "lb": [0, 0]})


def _pack_info(asg, Lv, Ix):
    gi = asg._capi.GridInfo()
    Lv = np.ascontiguousarray(Lv, dtype=np.int64)
    Ix = np.ascontiguousarray(Ix, dtype=np.int64)
    rc = asg._capi.load().asg_pack_info(Lv.shape[1], len(Lv), Lv.ctypes.data_as(asg._capi.i64p),
                                        Ix.ctypes.data_as(asg._capi.i64p), Lv.shape[1], 1, C.byref(gi))
    return rc, gi.as_dict(), asg._capi.load().asg_last_error().decode()


def test_packer_trie_sizes(asg):
    G = asg.SparseGridInterpolant(3)
    asg.rsg(G, 7, 5)
    rc, info, _ = _pack_info(asg, G.levels, G.indices)
    assert rc == 0 and info["canonical"] == 1 and info["subspaces"] == 72 and info["coef_slots"] == 737
    assert info["max_level"] == 5 and info["max_depth"] == 7 and info["hash_slots"] == 0 and info["regular"] == 1
    # drop one node / one whole subspace: still canonical, no longer a regular sparse grid
    keep = np.ones(len(G.levels), bool)
    keep[100] = False
    rc, info2, _ = _pack_info(asg, G.levels[keep], G.indices[keep])
    assert rc == 0 and info2["canonical"] == 1 and info2["regular"] == 0
    keep = ~np.all(G.levels == G.levels[-1], axis=1)
    rc, info3, _ = _pack_info(asg, G.levels[keep], G.indices[keep])
    assert rc == 0 and info3["canonical"] == 1 and info3["regular"] == 0 and info3["subspaces"] == 71
    # anisotropic caps are regular too
    A = asg.SparseGridInterpolant(4)
    asg.rsg(A, 6, (4, 5, 6, 3))
    rc, info4, _ = _pack_info(asg, A.levels, A.indices)
    assert rc == 0 and info4["regular"] == 1
    G = asg.SparseGridInterpolant(10)
    asg.rsg(G, 8, 8)
    rc, info, _ = _pack_info(asg, G.levels, G.indices)
    # 652 065 nodes, C(17,10) = 19 448 subspaces
    assert rc == 0 and info["N"] == 652065 and info["subspaces"] == 19448 and info["dense_subspaces"] == 19448
    assert info["trie_nodes"] == 43757 and info["coef_slots"] == 652065
    # sparse subspaces go to hash tables
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "cfg2_nodes.npz"))
    rc, info, _ = _pack_info(asg, g["key"][:, :2], g["key"][:, 2:])
    assert rc == 0 and info["canonical"] == 1 and info["subspaces"] == 148 and info["hash_slots"] > 0
    assert info["regular"] == 0


def test_packer_rejects_and_demotes(asg):
    E = asg._capi
    rc, _, msg = _pack_info(asg, [[1, 1], [2, 1]], [[1, 1], [1, 1]])
    assert rc == E.ASG_EINVALID_NODE and "(2, 1)" in msg
    rc, _, _ = _pack_info(asg, [[0, 1]], [[1, 1]])
    assert rc == E.ASG_EINVALID_NODE
    rc, _, _ = _pack_info(asg, [[63]], [[1]])
    assert rc == E.ASG_EUNSUPPORTED
    for Lv, Ix, why in [([[3, 1], [3, 1]], [[1, 1], [1, 1]], "duplicate"), ([[3, 1]], [[5, 1]], "range"),
                        ([[3, 1]], [[2, 1]], "range"), ([[1, 1]], [[3, 1]], "level-1"), ([[33]], [[1]], "32"),
                        ([[20, 20]], [[1, 1]], "32 bits")]:
        rc, info, msg = _pack_info(asg, Lv, Ix)
        assert rc == 0 and info["canonical"] == 0 and info["path"] == E.ASG_PATH_SWEEP and why in msg, (Lv, msg)
    rc, info, msg = _pack_info(asg, np.ones((1, 13), np.int64), np.ones((1, 13), np.int64))
    assert rc == 0 and info["canonical"] == 0 and "ASG_MAX_DIM_FAST" in msg


def test_packer_threads_report_like_the_serial_scan(asg):
    """Large grids are scanned and sorted by several host threads (asg_pack.cpp): errors and demotions must be the ones
    of the FIRST offending node in input order, whichever thread meets it, and the sort must give the serial order
    (same subspace / trie counts; the BLOCK layout's packer follows the device record stream on every pack)."""
    E = asg._capi
    G = asg.SparseGridInterpolant(8)
    asg.rsg(G, 7, 7)
    N = len(G.levels)
    assert N >= 1 << 15                                   # above the serial threshold
    rc, info, _ = _pack_info(asg, G.levels, G.indices)
    assert rc == 0 and info["canonical"] == 1 and info["regular"] == 1 and info["block"] == 1
    assert info["subspaces"] == 3003 and info["coef_slots"] == N          # C(14, 8) level vectors, every node once
    rng = np.random.default_rng(3)
    p = rng.permutation(N)                                # input order must not matter
    rc, info2, _ = _pack_info(asg, G.levels[p], G.indices[p])
    assert rc == 0 and {k: info2[k] for k in info if k != "generation"} == {k: info[k] for k in info if k != "generation"}
    # two invalid pairs, far apart: the first one in input order is reported
    Lv, Ix = G.levels.copy(), G.indices.copy()
    k2 = int(np.flatnonzero(Lv[:, 3] == 2)[-1])
    k1 = int(np.flatnonzero(Lv[:, 5] == 2)[0])
    assert k1 < k2
    Ix[k2, 3] = 1
    Ix[k1, 5] = 1
    rc, _, msg = _pack_info(asg, Lv, Ix)
    assert rc == E.ASG_EINVALID_NODE and f"at node {k1}, dim 5" in msg, msg
    # two demotions of different kinds: the reason of the first node
    Lv, Ix = G.levels.copy(), G.indices.copy()
    a = int(np.flatnonzero(Lv[:, 0] == 3)[0])
    b = int(np.flatnonzero(Lv[:, 1] == 1)[-1])
    assert a < b
    Ix[a, 0] = 2                                          # even index at level 3
    Ix[b, 1] = 3                                          # level-1 index != 1
    rc, info3, msg = _pack_info(asg, Lv, Ix)
    assert rc == 0 and info3["canonical"] == 0 and "range" in msg, msg
    Ix[a, 0] = 1
    rc, info3, msg = _pack_info(asg, Lv, Ix)
    assert rc == 0 and info3["canonical"] == 0 and "level-1" in msg, msg


def test_math_exports_host_helpers(asg, oracle):
    """The reference exports these next to phi (src/math.jl:9-22); the mirror's versions against the oracle's
    restatement and the docstring known answers (src/math.jl:513-514, 626-627, 997-999)."""
    assert [asg.fraction2li(k, 4) for k in range(5)] == [(2, 0), (3, 1), (1, 1), (3, 3), (2, 2)]
    for l in range(1, 8):
        assert tuple(map(list, asg.allnodes1d(l))) == tuple(map(list, oracle.allnodes1d(l)))
    for l, i in [(3, 1), (3, 3), (2, 0), (2, 2), (4, 5), (5, 7), (9, 201), (1, 1)]:
        L, I = asg.parent([1, l, 1], [1, i, 1], 2)
        assert (int(L[1]), int(I[1])) == oracle.parent1(l, i)
    with pytest.raises(ValueError):
        asg.parent([2], [1], 1)
    assert asg.isboundary(2, 0) and asg.isboundary(2, 2) and not asg.isboundary(3, 1) and not asg.isboundary(1, 1)
    assert [asg.boundaryflag(2, 0), asg.boundaryflag(2, 2), asg.boundaryflag(5, 3)] == [-1, 1, 0]
    with pytest.raises(ValueError):
        asg.boundaryflag(2, 1)
    L, I = asg.perturb([1, 2, 3], [1, 0, 3], 2, 4, 5)
    assert L.tolist() == [1, 4, 3] and I.tolist() == [1, 5, 3]
    assert asg.get_x(1, 1) == 0.5 and asg.get_x(2, 0) == 0.0 and asg.get_x(2, 2) == 1.0 and asg.get_x(4, 5) == 0.625
    assert np.array_equal(asg.scale(np.array([1.5, 2.0]), [1.0, 1.0], [2.0, 3.0]), [0.5, 0.5])
    assert asg.ghost_distance(4) == 0.125
    # ghost neighbours on the full level-l0 grid: walk the sorted node list of allnodes1d(l0)
    for l0 in (1, 2, 3, 5):
        Ls, Is = asg.allnodes1d(l0)
        nodes = list(zip(Ls, Is)) if l0 != 2 else [(2, 0), (1, 1), (2, 2)]   # allnodes1d(2) lists the root first (math.jl:626-627)
        for k, (l, i) in enumerate(nodes):
            for side, step in (("left", -1), ("right", 1)):
                L, I = asg.ghost_neighbor([l], [i], 1, l0, side)
                want = nodes[k + step] if 0 <= k + step < len(nodes) else (0, -1)
                assert (int(L[0]), int(I[0])) == want, (l0, l, i, side)
    with pytest.raises(ValueError):
        asg.ghost_neighbor([4], [3], 1, 3)      # node finer than the grid
