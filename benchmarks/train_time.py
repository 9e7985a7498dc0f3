"""cfg3 train! (D = 10, 652 065 nodes) and cfg4 train! + adapt! timed a few times in one process (A/B of library builds:
ASG_CUDA_LIB=<other libasg_cuda.so>)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import asg_b200 as asg

f = lambda X: np.sum(np.sin(2 * np.pi * X), axis=1)
asg.init(0)
for rep in range(3):
    G = asg.SparseGridInterpolant(10)
    t0 = time.perf_counter()
    asg.train(G, f, 8, 8, vectorized=True)
    t3 = time.perf_counter() - t0
    H = asg.SparseGridInterpolant(6)
    t0 = time.perf_counter()
    asg.train(H, f, 7, tuple(range(6, 12)), vectorized=True)
    t4 = time.perf_counter() - t0
    t0 = time.perf_counter()
    asg.adapt(H, f, 13, tol=3e-3, verbose=False, vectorized=True)
    t4a = time.perf_counter() - t0
    print(json.dumps({"rep": rep, "cfg3_train_s": t3, "cfg4_train_s": t4, "cfg4_adapt_s": t4a, "nodes": [len(G), len(H)],
                      "lib": os.environ.get("ASG_CUDA_LIB", "default")}), flush=True)
