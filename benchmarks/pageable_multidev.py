"""Pageable host arrays (row- and column-major) through the in-process multi-device path: 1 device vs 2, identical results."""
import sys, time, json
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import asg_b200 as asg
import bench
rng = np.random.default_rng(5)
M = 30_000_000
Xp = rng.random((M, 10))
Xf = np.asfortranarray(Xp)
ref = None
for devs in ([0], [0, 1]):
    asg.init_devices(devs)
    G = bench.build_grid_arrays()
    o1 = np.empty(M); o2 = np.empty(M)
    G.evaluate(Xp, out=o1)
    t0 = time.perf_counter(); G.evaluate(Xp, out=o1); t1 = time.perf_counter() - t0
    t0 = time.perf_counter(); G.evaluate(Xf, out=o2); t2 = time.perf_counter() - t0
    if ref is None: ref = o1.copy()
    print(json.dumps({"devices": len(devs), "points": M, "pageable_rowmajor_ms": t1 * 1e3, "pageable_colmajor_ms": t2 * 1e3,
                      "identical": bool(np.array_equal(o1, ref) and np.array_equal(o2, ref))}), flush=True)
    del G
    asg.shutdown()
