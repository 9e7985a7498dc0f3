"""Regenerates tests/golden/*.npz.  Run from the repo root:  python tests/golden/make_golden.py

The reference ships no golden vectors and Julia is not available, so the
committed vectors are of two kinds:
  * ANALYTIC: computed here with exact rational arithmetic (fractions.Fraction) straight from the
    formulas in the reference text (src/math.jl:332-357, 464-476) -- independent of oracle/;
  * ORACLE SNAPSHOTS: outputs of oracle/ on seeded inputs, committed so that the GPU tests can also
    run against fixed files and so that a change of the oracle is noticed.
"""
import os
import sys
from fractions import Fraction as Fr

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def phi_exact(x: Fr, l: int, i: int) -> Fr:
    """src/math.jl:345-357 in exact arithmetic."""
    if l > 2:
        t = x * 2 ** (l - 1) - i
        return 1 - abs(t) if -1 <= t <= 1 else Fr(0)
    if l == 2 and i == 0:
        return 1 - 2 * x if 0 <= x <= Fr(1, 2) else Fr(0)
    if l == 2 and i == 2:
        return 2 * x - 1 if Fr(1, 2) <= x <= 1 else Fr(0)
    if l == 1:
        return Fr(1) if 0 <= x <= 1 else Fr(0)
    raise ValueError


def main():
    rng = np.random.default_rng(20250101)
    # ---- analytic 1: phi on dyadic arguments (exactly representable in and out) -----------------
    rows = []
    for l in range(1, 12):
        idx = [1] if l == 1 else ([0, 2] if l == 2 else list(range(1, 2 ** (l - 1), 2)))
        for i in rng.choice(idx, size=min(len(idx), 6), replace=False):
            for _ in range(12):
                x = Fr(int(rng.integers(-64, 4096 + 64)), 4096)
                rows.append((float(x), l, int(i), float(phi_exact(x, l, int(i)))))
    rows = np.array(rows)
    np.savez_compressed(os.path.join(HERE, "phi_dyadic.npz"), x=rows[:, 0], l=rows[:, 1].astype(np.int64),
                        i=rows[:, 2].astype(np.int64), phi=rows[:, 3])

    # ---- analytic 2: f = prod 4x(1-x), D = 3, train!(6, 6): closed-form surpluses ------------------
    from oracle import oracle as o
    G = o.OracleGrid(3)
    G.rsg(6, 6)
    s1 = lambda l: Fr(1) if l == 1 else (Fr(-1) if l == 2 else Fr(4) ** (2 - l))
    coefs = np.array([float(s1(a) * s1(b) * s1(c)) for a, b, c in G.levels])
    # exact values of the interpolant at dyadic points from the exact surpluses
    pts = [[Fr(int(v), 64) for v in rng.integers(0, 65, size=3)] for _ in range(200)]
    vals = []
    for p in pts:
        acc = Fr(0)
        for (L, I, c) in zip(G.levels, G.indices, coefs):
            t = Fr(c)
            for d in range(3):
                t *= phi_exact(p[d], int(L[d]), int(I[d]))
                if t == 0:
                    break
            acc += t
        vals.append(float(acc))
    np.savez_compressed(os.path.join(HERE, "parabola_d3.npz"), levels=G.levels, indices=G.indices, coefs=coefs,
                        X=np.array([[float(v) for v in p] for p in pts]), y=np.array(vals))

    # ---- snapshot 1: cfg1 (README D=3 RSG train!(G, sin-sum, 7, 5)) --------------------------------
    f = lambda x: float(np.sum(np.sin(2 * np.pi * x)))
    G1 = o.OracleGrid(3)
    G1.train(f, 7, 5)
    X = rng.random((2000, 3))
    y, ab = G1.evaluate(X, return_abssum=True)
    y4 = G1.evaluate(X, max_depth=4)
    np.savez_compressed(os.path.join(HERE, "cfg1_eval.npz"), levels=G1.levels, indices=G1.indices, depths=G1.depths,
                        coefs=G1.coefs, fvals=G1.fvals, X=X, y=y, abssum=ab, y_depth4=y4)

    # ---- snapshot 2: cfg2 node set (train!(4,(3,8)) + adapt!(20, 1e-5) on -1/(0.1+sum x)) ----------
    h = lambda x: -1.0 / (0.1 + float(np.sum(x)))
    G2 = o.OracleGrid(2)
    G2.train(h, 4, (3, 8))
    log = []
    G2.adapt(h, 20, tol=1e-5, log=log)
    key = np.concatenate([G2.levels, G2.indices], axis=1)
    order = np.lexsort(key.T[::-1])
    np.savez_compressed(os.path.join(HERE, "cfg2_nodes.npz"), key=key[order].astype(np.int32), coefs=G2.coefs[order],
                        added=np.array([l["added"] for l in log]))
    print("wrote golden files; cfg2 nodes:", len(G2), "added per level:", [l["added"] for l in log])


if __name__ == "__main__":
    main()
