"""Host-buffer evaluation from PAGEABLE memory (numpy arrays, Julia Arrays) against pinned buffers: cfg2 (D = 2, 1e7
points) and cfg3 (D = 10, 2e7 points). ASG_STAGE_HOST=0 leaves pageable buffers to the driver's staging (A/B)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import asg_b200 as asg
import bench

asg.init(0)
rng = np.random.default_rng(1)


def timed(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


u = lambda X: -1.0 / (0.1 + np.sum(X, axis=1))
A = asg.SparseGridInterpolant(2)
asg.train(A, u, 4, (3, 8), vectorized=True)
asg.adapt(A, u, 20, tol=1e-5, verbose=False, vectorized=True)
G = bench.build_grid_arrays()
for name, H, M in (("cfg2", A, 10_000_000), ("cfg3", G, 20_000_000)):
    D = H.levels.shape[1]
    Xp = rng.random((M, D))                                   # pageable
    op = np.empty(M)
    Xh = torch.empty((M, D), dtype=torch.float64, pin_memory=True)
    Xh.numpy()[:] = Xp
    oh = torch.empty(M, dtype=torch.float64, pin_memory=True)
    tp = timed(lambda: H.evaluate(Xp, out=op))
    th = timed(lambda: H.evaluate(Xh.numpy(), out=oh.numpy()))
    same = bool(np.array_equal(op, oh.numpy()))
    Xf = np.asfortranarray(Xp)                                # column-major: Julia's M x D Matrix
    of = np.empty(M)
    tf = timed(lambda: H.evaluate(Xf, out=of))
    same = same and bool(np.array_equal(of, op))
    print(json.dumps({"config": name, "points": M, "pageable_ms": tp * 1e3, "pageable_colmajor_ms": tf * 1e3,
                      "pinned_ms": th * 1e3, "pageable_points_per_s": M / tp, "pinned_points_per_s": M / th,
                      "identical": same, "stage_host": os.environ.get("ASG_STAGE_HOST", "1")}), flush=True)
