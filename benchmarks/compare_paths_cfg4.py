"""Cross-check of the three evaluation engines (BLOCK walk, flat stream, sweep) on the grids adapt! grows for
BASELINE configs[3] (D = 6, 15 089 -> 19 929 nodes): max |difference| against the sweep engine per refinement depth.
Run on a B200: python benchmarks/compare_paths_cfg4.py"""
import sys, numpy as np
sys.path.insert(0, '.')
import asg_b200 as asg
asg.init(0)
def sinsum_v(X): return np.sum(np.sin(2*np.pi*X), axis=1)
G = asg.SparseGridInterpolant(6)
asg.train(G, sinsum_v, 7, tuple(range(6, 12)), vectorized=True)
G.set_path("stream")
for md in (8, 9, 10, 11, 12, 13):
    asg.adapt(G, sinsum_v, md, tol=3e-3, verbose=False, vectorized=True)
    rng = np.random.default_rng(md)
    X = np.vstack([rng.random((20000, 6)), rng.integers(0, 4097, size=(20000, 6)) / 4096.0])
    res = {}
    for path in ("stream", "sweep", "block"):
        try:
            G.set_path(path)
        except Exception as e:
            print(md, path, "unavailable:", e); continue
        res[path] = G.evaluate(X)
    G.set_path("stream")
    info = G.info()
    print(md, len(G), {k: info[k] for k in ("block", "block_slots", "block_records", "block_stack", "block_tiles")},
          {p: float(np.max(np.abs(res[p] - res["sweep"]))) for p in res})
    if "block" in res:
        bad = np.flatnonzero(np.abs(res["block"] - res["sweep"]) > 1e-12)
        print("   bad points:", len(bad), X[bad[:3]] if len(bad) else "")
