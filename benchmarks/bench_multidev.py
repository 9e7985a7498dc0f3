#!/usr/bin/env python
"""Several GPUs behind the C ABI, ONE process (asg_init_devices): end-to-end throughput of G.evaluate(X) on the cfg3
grid with pinned host buffers, for 1, 2, 4, ... all visible devices. No torch.distributed, no NCCL: the library splits
the host range over its device replicas (one host thread + one copy/compute pipeline per device).
    python benchmarks/bench_multidev.py [--points 1e8] [--steps 3]   ->  one JSON line per device count"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import asg_b200 as asg  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=float, default=1e8)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--devices", type=str, default="")
    a = ap.parse_args()
    M = int(a.points)
    ng = torch.cuda.device_count()
    counts = [int(c) for c in a.devices.split(",")] if a.devices else [n for n in (1, 2, 4, 8) if n <= ng]
    rng = np.random.default_rng(42)
    Xh = torch.empty((M, bench.D), dtype=torch.float64, pin_memory=True)
    Xn = Xh.numpy()
    for lo in range(0, M, 1 << 22):
        Xn[lo:lo + (1 << 22)] = rng.random((min(1 << 22, M - lo), bench.D))
    oh = torch.empty(M, dtype=torch.float64, pin_memory=True)
    on = oh.numpy()
    ref = None
    for n in counts:
        asg.init_devices(list(range(n)))
        G = bench.build_grid_arrays()
        t0 = time.perf_counter()
        G._ensure()
        t_up = time.perf_counter() - t0
        G.evaluate(Xn, out=on)  # warm-up: allocates the pipeline buffers of every replica
        t0 = time.perf_counter()
        for _ in range(a.steps):
            G.evaluate(Xn, out=on)
        dt = (time.perf_counter() - t0) / a.steps
        chk = float(on.sum())
        if ref is None:
            ref = on.copy()
        same = bool(np.array_equal(ref, on))
        print(json.dumps({"devices": n, "points": M, "ms_per_call": dt * 1e3, "points_per_s": M / dt,
                          "upload_all_replicas_s": t_up, "checksum": chk, "bit_identical_to_first": same}), flush=True)
        del G
        asg.shutdown()


if __name__ == "__main__":
    main()
