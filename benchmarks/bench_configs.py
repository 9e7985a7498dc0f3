#!/usr/bin/env python
"""Secondary measurements: the other BASELINE.json configs (parity-test cases, not the headline line).
One JSON line per config; run on a B200:  python benchmarks/bench_configs.py > gpurun_out/configs.jsonl

cfg1  README D=3 RSG train!(7,5), evaluate 1e5 points            (+ oracle check on 2000 points)
cfg2  D=2 train!(4,(3,8)) + adapt!(20, 1e-5), evaluate 1e7 points (hashed subspaces, generic groups)
cfg4  D=6 train!(7,(6..11)) + adapt!(13, 3e-3): whole loop with the GPU surplus step per level
cfg5  D=4 RSG(10,10): basis(X,G) sparse (count/fill/sort) and dense (HBM-write bound) for 5e4 points
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import asg_b200 as asg  # noqa: E402

sinsum = lambda x: float(np.sum(np.sin(2 * np.pi * x)))
sinsum_v = lambda X: np.sum(np.sin(2 * np.pi * X), axis=1)
utility = lambda x: -1.0 / (0.1 + float(np.sum(x)))
utility_v = lambda X: -1.0 / (0.1 + np.sum(X, axis=1))


def timed(fn, reps=3):
    fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t.append(time.perf_counter() - t0)
    return min(t)


def device_eval_ms(lib, G, X, reps=5):
    """points resident in HBM: asg_eval_device timed with CUDA events inside the library (median of `reps`)"""
    import torch
    dX = torch.from_numpy(np.ascontiguousarray(X)).cuda()
    out = torch.empty(len(X), dtype=torch.float64, device="cuda")
    h = G._ensure()
    st = torch.cuda.current_stream().cuda_stream
    ms = C.c_double()
    ts = []
    for k in range(reps + 2):
        asg._capi.check(lib.asg_eval_device(h, C.c_void_p(dX.data_ptr()), len(X), X.shape[1], 1, -1, C.c_void_p(out.data_ptr()),
                                            C.c_void_p(st)))
        asg._capi.check(lib.asg_last_kernel_ms(C.byref(ms)))
        if k >= 2:
            ts.append(ms.value)
    return float(np.median(ts))


def main():
    import torch
    asg.init(0)
    lib = asg._capi.load()
    rng = np.random.default_rng(42)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass

    # ---- cfg1 -------------------------------------------------------------------------------------
    G = asg.SparseGridInterpolant(3)
    t0 = time.perf_counter()
    asg.train(G, sinsum_v, 7, 5, vectorized=True)
    t_train = time.perf_counter() - t0
    X = rng.random((100000, 3))
    dt = timed(lambda: G.evaluate(X))
    from oracle import oracle as o
    OG = o.OracleGrid(3)
    OG.levels, OG.indices, OG.depths, OG.coefs, OG.fvals = G.levels, G.indices, G.depths, G.coefs, G.fvals
    err = float(np.max(np.abs(G.evaluate(X[:2000]) - OG.evaluate(X[:2000]))))
    print(json.dumps({"config": "cfg1", "nodes": len(G), "train_s": t_train, "points": len(X), "eval_ms": dt * 1e3,
                      "points_per_s": len(X) / dt, "max_abs_err_vs_oracle": err, "info": G.info()}), flush=True)

    # ---- cfg2 -------------------------------------------------------------------------------------
    G = asg.SparseGridInterpolant(2)
    t0 = time.perf_counter()
    asg.train(G, utility_v, 4, (3, 8), vectorized=True)
    asg.adapt(G, utility_v, 20, tol=1e-5, verbose=False, vectorized=True)
    t_build = time.perf_counter() - t0
    X = rng.random((10_000_000, 2))
    dt = timed(lambda: G.evaluate(X), reps=2)
    dms = device_eval_ms(lib, G, X)
    print(json.dumps({"config": "cfg2", "nodes": len(G), "train_adapt_s": t_build, "points": len(X),
                      "eval_ms_host_buffers_pageable": dt * 1e3, "points_per_s_host_buffers": len(X) / dt,
                      "eval_ms_device_resident": dms, "points_per_s_device_resident": len(X) / (dms * 1e-3),
                      "info": G.info()}), flush=True)

    # ---- cfg4 -------------------------------------------------------------------------------------
    G = asg.SparseGridInterpolant(6)
    l0 = asg.launch_count()
    t0 = time.perf_counter()
    asg.train(G, sinsum_v, 7, tuple(range(6, 12)), vectorized=True)
    t_train = time.perf_counter() - t0
    n_rsg = len(G)
    log = []
    t0 = time.perf_counter()
    asg.adapt(G, sinsum_v, 13, tol=3e-3, verbose=False, vectorized=True, log=log)
    t_adapt = time.perf_counter() - t0
    print(json.dumps({"config": "cfg4", "rsg_nodes": n_rsg, "nodes": len(G), "train_s": t_train, "adapt_s": t_adapt,
                      "levels": log, "kernel_launches": asg.launch_count() - l0, "info": G.info()}), flush=True)

    # ---- cfg3 train!: the full D=10 grid trained on the GPU (8 fused surplus calls, 8.8e10 node*point pairs) ----
    G = asg.SparseGridInterpolant(10)
    t0 = time.perf_counter()
    asg.rsg(G, 8, 8)
    t_rsg = time.perf_counter() - t0
    t0 = time.perf_counter()
    asg.train(G, sinsum_v, 8, 8, vectorized=True)
    t_train = time.perf_counter() - t0           # includes a second rsg! like the reference's train!
    Xn = asg.stack(G)[::97]
    err = float(np.max(np.abs(G.evaluate(Xn) - G.fvals[::97])))
    print(json.dumps({"config": "cfg3-train", "nodes": len(G), "rsg_s": t_rsg, "train_total_s": t_train,
                      "interpolation_err_at_nodes": err, "pairs": 8.8e10}), flush=True)

    # ---- device-format persistence of the trained cfg3 grid -----------------------------------
    t0 = time.perf_counter()
    G.sync()
    t_upload = time.perf_counter() - t0                     # validate + sort + pack + upload
    t0 = time.perf_counter()
    dat = asg.serialize(G, packed=True)
    t_export = time.perf_counter() - t0
    t0 = time.perf_counter()
    H = asg.deserialize(dat)
    t_import = time.perf_counter() - t0                     # checksum + upload
    same = bool(np.array_equal(G.evaluate(Xn), H.evaluate(Xn)))
    print(json.dumps({"config": "cfg3-persist", "nodes": len(G), "upload_s": t_upload, "export_s": t_export,
                      "import_s": t_import, "blob_bytes": int(dat["packed"].size), "identical_results": same}), flush=True)
    del H, dat

    # ---- cfg5 -------------------------------------------------------------------------------------
    G = asg.SparseGridInterpolant(4)
    asg.rsg(G, 10, 10)
    G.coefs = rng.standard_normal(len(G))
    N = len(G)
    M = 50000
    X = rng.random((M, 4))
    Xe = rng.random((10_000_000, 4))
    dms5 = device_eval_ms(lib, G, Xe)
    del Xe
    # where the sparse basis spends its time: the two device passes vs. the host side (pageable D2H, scipy)
    import scipy.sparse as sps
    h5 = G._ensure()
    Xc = np.ascontiguousarray(X)
    vp = lambda a: C.c_void_p(a.ctypes.data)
    t0 = time.perf_counter()
    nnz_row = np.zeros(M, dtype=np.int64)
    asg._capi.check(lib.asg_basis_count(h5, vp(Xc), M, 4, 1, 1e-16, 1, nnz_row.ctypes.data_as(asg._capi.i64p)))
    t_count = time.perf_counter() - t0
    ptr = np.zeros(M + 1, dtype=np.int64)
    np.cumsum(nnz_row, out=ptr[1:])
    col = np.empty(int(ptr[-1]), dtype=np.int64)
    val = np.empty(int(ptr[-1]), dtype=np.float64)
    cptr = np.empty(N + 1, dtype=np.int64)
    t0 = time.perf_counter()
    asg._capi.check(lib.asg_basis_fill_csc(h5, vp(Xc), M, 4, 1, 1e-16, 1, ptr.ctypes.data_as(asg._capi.i64p),
                                           cptr.ctypes.data_as(asg._capi.i64p), col.ctypes.data_as(asg._capi.i64p),
                                           val.ctypes.data_as(asg._capi.f64p)))
    t_fill = time.perf_counter() - t0
    t0 = time.perf_counter()
    sps.csc_matrix((val, col, cptr), shape=(M, N))
    t_scipy = time.perf_counter() - t0
    del col, val
    t_sparse = timed(lambda: asg.basis(X, G), reps=2)
    t_sparse_csr = timed(lambda: asg.basis(X, G, format="csr"), reps=2)   # without the host-side CSR -> CSC transposition
    B = asg.basis(X, G)
    y = G.evaluate(X)
    err = float(np.max(np.abs(B @ G.coefs - y)))
    # dense variant into a device buffer: HBM-write bound, 8*N bytes per point
    dX = torch.from_numpy(X).cuda()
    out = torch.empty((M, N), dtype=torch.float64, device="cuda")
    h = G._ensure()
    st = torch.cuda.current_stream().cuda_stream

    def dense():
        asg._capi.check(lib.asg_basis_dense_device(h, C.c_void_p(dX.data_ptr()), M, 4, 1, 1e-16, 1,
                                                   C.c_void_p(out.data_ptr()), N, 1, C.c_void_p(st)))
        torch.cuda.synchronize()
    for _ in range(3):
        dense()
    ms = C.c_double()
    ts = []
    for _ in range(5):
        dense()
        asg._capi.check(lib.asg_last_kernel_ms(C.byref(ms)))
        ts.append(ms.value)
    k_ms = float(np.median(ts))
    gbs = M * N * 8 / (k_ms * 1e-3) / 1e9
    nnz_dense = int(torch.count_nonzero(out).item())
    print(json.dumps({"config": "cfg5", "nodes": N, "points": M, "eval_1e7_points_device_resident_ms": dms5,
                      "eval_points_per_s_device_resident": 1e7 / (dms5 * 1e-3),
                      "sparse_breakdown_s": {"count_call": t_count, "fill_csc_call_incl_576MB_pageable_D2H": t_fill,
                                             "scipy_csc_matrix": t_scipy},
                      "sparse_basis_s": t_sparse, "sparse_basis_csr_s": t_sparse_csr, "nnz": int(B.nnz),
                      "nnz_per_row": B.nnz / M, "B_times_C_vs_eval_max_abs": err, "dense_ms": k_ms,
                      "dense_bytes": M * N * 8, "dense_write_GBps": gbs, "hbm_peak_GBps": peaks.get("hbm_gbs"),
                      "dense_frac_of_hbm_peak": gbs / peaks["hbm_gbs"] if "hbm_gbs" in peaks else None,
                      "dense_nnz_matches_sparse": nnz_dense == int(B.nnz)}), flush=True)


if __name__ == "__main__":
    main()
