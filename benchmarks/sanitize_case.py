"""One small scenario through every kernel of the library, to be run under compute-sanitizer:

    compute-sanitizer --tool memcheck  python benchmarks/sanitize_case.py
    compute-sanitizer --tool racecheck python benchmarks/sanitize_case.py      (shared-memory hazards, TMA tiles included)
    compute-sanitizer --tool synccheck python benchmarks/sanitize_case.py
    compute-sanitizer --tool initcheck python benchmarks/sanitize_case.py

Sizes are chosen so that each engine takes the code path it takes at benchmark size (two points per thread, cell-key
ordering, PAIR mode with its follow-up launch, WIDE suffix budgets, depth masks, the delta sweep after an append) while
a tool that slows kernels down 100x still finishes in minutes. Engines are compared with each other, not with the
oracle (which only tests/, smoke() and bench.py may use). ASG_SANITIZE_SMALL=1 shrinks the batches further (racecheck).
Run plainly it is a cross-check of the engines (compute-sanitizer is closed on the round-2 GPU pool, exit code 86).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import asg_b200 as asg

small = os.environ.get("ASG_SANITIZE_SMALL", "0") == "1"
asg.init(0)
rng = np.random.default_rng(0)
f = lambda X: np.sum(np.sin(2 * np.pi * X), axis=1)


def agree(a, b, what, tol=1e-12):
    err = float(np.max(np.abs(np.asarray(a) - np.asarray(b))))
    assert err <= tol, f"{what}: engines differ by {err:.3e}"
    return err


# 1. regular D = 4 grid: stream walk one / two points per thread, masked, surplus, dense + sparse basis
G = asg.SparseGridInterpolant(4)
asg.train(G, f, 6, (4, 5, 6, 3), vectorized=True)
X = rng.random((40000 if small else 330000, 4))
G.set_path("stream")
y = G.evaluate(X)
y4 = G.evaluate(X, max_depth=4)
s = G.surplus(X[:1000], y[:1000])
print("regular", len(G), float(np.abs(s).max()), G.info()["regular"], flush=True)
G.set_path("block")
agree(G.evaluate(X), y, "block vs stream, D = 4")
agree(G.evaluate(X, max_depth=4), y4, "block vs stream, D = 4, depth <= 4")
agree(G.evaluate(X[:700]), y[:700], "block, one point per thread")
for p in ("trie", "sweep"):
    G.set_path(p)
    agree(G.evaluate(X[:5000]), y[:5000], f"{p} vs stream")
G.set_path("auto")
Bd = asg.basis(X[:100], G, dense=True)
Bs = asg.basis(X[:100], G)
Bc = asg.basis(X[:100], G, format="csr")
agree(Bs.toarray(), Bd, "sparse vs dense basis", 0.0)
agree(Bc.toarray(), Bd, "csr vs dense basis", 0.0)
E = asg.dehierarchization_matrix(G)
c_train = G.coefs.copy()
agree(asg.hierarchize(G), c_train, "hierarchize vs train! coefficients", 1e-10)
print("basis", Bd.shape, Bs.nnz, E.nnz, flush=True)

# 2. D = 6 depth-6 grid (924 subspaces: cell-key ordering), BLOCK walk with PAIR mode forced on + the follow-up launch
H = asg.SparseGridInterpolant(6)
asg.rsg(H, 6, 6)
H.coefs = rng.standard_normal(len(H))
M = 80000 if small else 200000          # above one wave of 148 x 512 threads: two points per thread
X6 = rng.random((M, 6))
X6[::101, 2] = 1.0
X6[::211, 4] = np.nan
X6[::307, 1] = -0.25
H.set_path("block")
y6 = H.evaluate(X6)
os.environ["ASG_BLOCK_PAIR"] = "1e-9"
yp = H.evaluate(X6)
ypm = H.evaluate(X6, max_depth=4)
del os.environ["ASG_BLOCK_PAIR"]
assert np.array_equal(yp, y6), "PAIR mode changes the sums"
assert np.array_equal(ypm, H.evaluate(X6, max_depth=4)), "PAIR mode changes the masked sums"
H.set_path("stream")
agree(H.evaluate(X6[:20000]), y6[:20000], "stream vs block, D = 6")
print("pair", len(H), H.info()["block_pair"], flush=True)

# 3. WIDE suffix budgets (levels 9, 10): D = 3 depth 9
W = asg.SparseGridInterpolant(3)
asg.rsg(W, 9, 9)
W.coefs = rng.standard_normal(len(W))
Xw = rng.random((30000 if small else 160000, 3))
W.set_path("block")
yw = W.evaluate(Xw)
W.set_path("stream")
agree(W.evaluate(Xw), yw, "stream vs wide block walk")
W.set_path("block")
agree(W.evaluate(Xw[:500], max_depth=7), (W.set_path("stream"), W.evaluate(Xw[:500], max_depth=7))[1], "wide, masked")
print("wide", len(W), flush=True)

# 4. adaptive grid: hashed subspaces, candidate generation, appended rows (delta sweep), persistence
A = asg.SparseGridInterpolant(2)
u = lambda X: -1.0 / (0.1 + np.sum(X, axis=1))
asg.train(A, u, 4, (3, 8), vectorized=True)
asg.adapt(A, u, 12, tol=1e-4, verbose=False, vectorized=True)
Xa = rng.random((30000 if small else 310000, 2))
ya = A.evaluate(Xa)
B = asg.basis(rng.random((200, 2)), A)
for p in ("trie", "sweep"):
    A.set_path(p)
    agree(A.evaluate(Xa[:5000]), ya[:5000], f"adaptive, {p}")
A.set_path("auto")
A2 = asg.deserialize(asg.serialize(A, packed=True))       # the export folds the appended rows into the packed layouts
ya2 = A.evaluate(Xa[:5000])
agree(ya2, ya[:5000], "packed + delta sweep vs consolidated")
agree(A2.evaluate(Xa[:5000]), ya2, "packed blob", 0.0)
z = asg.extrapolate(A, rng.random((300, 2)) * 1.4 - 0.2)
print("adaptive", len(A), B.nnz, A.info()["packs"], A.info()["delta_nodes"], float(np.abs(z).max()), flush=True)

# 5. the exported basis function
p = asg.phi(rng.random((1000, 1)), rng.integers(3, 12, size=(1000, 1)), np.ones((1000, 1), dtype=np.int64))
print("ok", float(np.sum(p)), flush=True)
