"""Differential fuzz of the evaluation engines on the GPU: random grids (dimension, depth, level caps, node subsets,
domains, adapt!-grown) evaluated by the BLOCK walk, the flat stream, the trie walk and the sweep engine (the sweep is
the reference's loop nest turned inside out and needs no layout at all) -- unmasked and depth-masked, small and large
batches. Prints one line per grid and fails on the first disagreement.   python benchmarks/fuzz_paths.py [n] [seed]"""
import sys

import numpy as np

sys.path.insert(0, ".")
import asg_b200 as asg

asg.init(0)
n_grids = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
kinds = {"block": 0, "noblock": 0, "split": 0, "deep": 0}
for t in range(n_grids):
    D = int(rng.integers(1, 13))
    depth = int(rng.integers(2, {1: 12, 2: 11, 3: 9, 4: 9}.get(D, 7 if D <= 7 else 5)))
    caps = tuple(int(v) for v in rng.integers(2, depth + 2, size=D)) if rng.random() < 0.5 else depth
    lb = np.zeros(D) if rng.random() < 0.4 else -rng.random(D) * 3
    ub = np.ones(D) if rng.random() < 0.4 else lb + 0.5 + rng.random(D) * 4
    G = asg.SparseGridInterpolant(D, lb=lb, ub=ub)
    asg.rsg(G, depth, caps)
    if len(G) > 400000:
        continue
    mode = rng.choice(["full", "subset", "adapt"])
    if mode == "subset":
        keep = rng.random(len(G)) < rng.choice([0.97, 0.8, 0.5, 0.2])
        keep[0] = True
        G.levels, G.indices, G.depths = G.levels[keep], G.indices[keep], G.depths[keep]
    G.coefs = rng.standard_normal(len(G.levels)) * 2.0 ** (-rng.random() * G.depths)
    G.fvals = np.zeros(len(G.levels))
    if mode == "adapt" and D <= 6:
        c = rng.random(D)
        f = lambda X: np.exp(-np.sum(((X - lb) / (ub - lb) - c) ** 2, axis=1) * 8.0)
        asg.train(G, f, depth, caps, vectorized=True)
        asg.adapt(G, f, depth + int(rng.integers(1, 4)), tol=10.0 ** -rng.integers(2, 5), verbose=False, vectorized=True)
    info = G.info()
    M = int(rng.choice([257, 3000, 40000, 1 << 20])) if len(G.levels) < 60000 else int(rng.choice([257, 3000, 40000]))
    u = rng.random((M, D))
    u[rng.random((M, D)) < 0.05] = rng.integers(0, 65, size=1) / 64.0          # knots
    u[::97, rng.integers(0, D)] = rng.choice([1.0, 0.0, 1.25, -0.5, np.nan])
    X = lb + (ub - lb) * u
    L = int(rng.integers(1, int(G.depths.max()) + 1))
    G.set_path("sweep")
    want, wantm = G.evaluate(X), G.evaluate(X, max_depth=L)
    scale = max(1.0, float(np.nanmax(np.abs(want)))) * max(1, len(G.levels)) ** 0.5
    paths = ["stream", "trie"] + (["block"] if info["block"] else []) if info["canonical"] else []
    worst = 0.0
    for p in paths:
        G.set_path(p)
        for got, ref, what in ((G.evaluate(X), want, "all"), (G.evaluate(X, max_depth=L), wantm, f"depth<={L}")):
            err = float(np.nanmax(np.abs(got - ref))) if len(ref) else 0.0
            nan_mismatch = bool(np.any(np.isnan(got) != np.isnan(ref)))
            worst = max(worst, err)
            if err > 1e-13 * scale or nan_mismatch:
                print(f"MISMATCH grid {t}: D={D} depth={depth} caps={caps} mode={mode} N={len(G.levels)} M={M} path={p} {what} "
                      f"err={err:.3e} info={info}")
                sys.exit(1)
    kinds["block" if info["block"] else "noblock"] += 1
    kinds["split"] += int(info["block"] and info["block_records"] > info["subspaces"] * 0 + 0 and info["block_tiles"] > 1)
    kinds["deep"] += int(info["block"] and info["block_stack"] > 2)
    print(f"grid {t}: D={D} depth={depth} mode={mode} N={len(G.levels)} M={M} block={info['block']} records={info['block_records']} "
          f"stack={info['block_stack']} max|diff|={worst:.2e}", flush=True)
print("all grids agree", kinds)
