#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json: "interpolant evals/sec and
node*point products/sec at 1/2/4/8 B200 vs CPU").

Workload (BASELINE.json configs[2], "cfg3"): D = 10 isotropic regular sparse grid, depth 8 / max level
8 (652 065 nodes, 19 448 subspaces), batched G(X) over 1e8 uniformly random fp64 points; the grid is
replicated, the points are split into contiguous shards over the ranks (total fixed -> "strong").
Synthetic data: coefficients N(0,1) * 2^-depth (seed 1234), points U[0,1)^10 (seed 42 + rank).

One "step" = one evaluation of all M points.
  value  : points/s with the points already resident in HBM (asg_eval_device), CUDA events, max over ranks
  e2e    : points/s through the public API G.evaluate(X) with pinned HOST buffers (H2D + D2H inside)
  roofline / cpu_baseline / clocks / gpu_launches: see DESIGN.md "Measurement"

  python bench.py [--gpus N] [--steps K] [--warmup W] [--points M] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D, DEPTH, MAXLEVEL = 10, 8, 8
WORKLOAD = "cfg3: D=10 isotropic RSG depth 8 / maxlevel 8 (652065 nodes), batched G(X)"


def build_grid_arrays():
    """Host enumeration of the node set (rsg!, src/train.jl:40-78) + synthetic coefficients."""
    import asg_b200 as asg
    G = asg.SparseGridInterpolant(D)
    asg.rsg(G, DEPTH, MAXLEVEL)
    rng = np.random.default_rng(1234)
    G.coefs = rng.standard_normal(len(G)) * 2.0 ** (-G.depths.astype(np.float64))
    return G


def build_oracle_grid():
    """The same grid built by the ORACLE's own enumeration (the reference arm maps nothing of the product)."""
    from oracle import oracle as o
    OG = o.OracleGrid(D)
    OG.rsg(DEPTH, MAXLEVEL)
    rng = np.random.default_rng(1234)
    OG.coefs = rng.standard_normal(len(OG)) * 2.0 ** (-OG.depths.astype(np.float64))
    return OG


def fp64_op_counts(levels):
    """FP64 work the stream kernel (asg_kernels_stream.cu, SUPER records) must do for ONE point, counted
    from the level vectors of `levels` (N x D). Dimensions 0..D-3 are walked: every trie node there with
    level > 1 costs DADD (int->double) + DFMA + DADD (1-|t|) + DMUL (prefix product) = 4 instr / 5 flop.
    Dimensions D-2, D-1 come from per-point tables: 1 DFMA per subspace with l_{D-1} > 1, 1 DFMA per prefix
    group with l_{D-2} > 1, 1 DFMA per level-(D-3) node (acc += P * inner); the tables themselves cost
    2 * 7 * 3 instr once per point. Returns (fp64 instructions, flops, subspaces)."""
    lv = np.unique(levels, axis=0)
    D = lv.shape[1]
    nontrivial = 0
    for k in range(1, max(D - 2, 0) + 1):
        pre = np.unique(lv[:, :k], axis=0)
        nontrivial += int(np.count_nonzero(pre[:, k - 1] > 1))
    leaves = len(lv)
    leaf_fma = int(np.count_nonzero(lv[:, -1] > 1))
    groups = np.unique(lv[:, :-1], axis=0) if D >= 2 else np.zeros((1, 0))
    group_fma = int(np.count_nonzero(groups[:, -1] > 1)) if D >= 2 else 0
    supers = len(np.unique(lv[:, :-2], axis=0)) if D >= 3 else 1
    fma = leaf_fma + group_fma + supers
    instr = 4 * nontrivial + fma + 42
    flops = 5 * nontrivial + 2 * fma + 2 * 7 * 4
    return instr, flops, leaves


def fp64_op_counts_block(levels, records):
    """FP64 work of the BLOCK walk (asg_kernels_block.cu) for ONE point, counted from the level vectors and the
    number of records the packer formed. Per record: stage DMUL (x*2^b) + DADD (int->double) + DFMA (t) + DADD (1-|t|)
    + DMUL (prefix product) and the record's DFMA into the accumulator = 6 instr / 8 flop. Every other trie node of
    dimensions 0..D-3 with level > 1 is compiled: DADD + DFMA + DADD (hat on the fly) + DFMA (Horner) = 4 instr /
    6 flop (plus DMUL x*2^b where the cell key is not in a register: dimensions above D-3). Per pole with l_{D-2} > 1
    and per subspace with l_{D-1} > 1: one DFMA. Register tables: 2 dims x 3 levels x 3 instr once per point.
    Returns (fp64 instructions, flops, subspaces)."""
    lv = np.unique(levels, axis=0)
    D = lv.shape[1]
    leaf_fma = int(np.count_nonzero(lv[:, -1] > 1))
    poles = np.unique(lv[:, :-1], axis=0)
    pole_fma = int(np.count_nonzero(poles[:, -1] > 1))
    nt_A = int(np.count_nonzero(np.unique(lv[:, :-2], axis=0)[:, -1] > 1))   # nodes of dimension D-3 with level > 1
    nt_pre = sum(int(np.count_nonzero(np.unique(lv[:, :k], axis=0)[:, -1] > 1)) for k in range(1, D - 2))
    compiled_pre = nt_pre - (records - 1)          # prefix nodes that are not a record's stage (deeper suffixes)
    instr = leaf_fma + pole_fma + 4 * nt_A + 5 * compiled_pre + 6 * records + 18
    flops = 2 * (leaf_fma + pole_fma) + 6 * nt_A + 7 * compiled_pre + 8 * records + 2 * 3 * 4
    return instr, flops, len(lv)


def pin_to_gpu_numa_node(index):
    """Multi-rank runs: bind this process (and the pinned host buffers it allocates afterwards) to the CPU cores
    NVML reports as local to its GPU, so that the eight ranks' H2D streams do not all cross the socket link."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return None
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.th.join(timeout=5)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_rate(Gh, seconds_target, threads=0, steps=1, warmup=0):
    """Reference-structured CPU path (oracle/asg_oracle.c: asgo_refpath_batch, the cost model of
    src/class.jl:635-642) on a bounded sample of the same workload. Returns points/s and a description."""
    from oracle import oracle as o
    OG = o.OracleGrid(D)
    OG.levels, OG.indices, OG.depths, OG.coefs, OG.fvals = Gh.levels, Gh.indices, Gh.depths, Gh.coefs, Gh.fvals
    cores = threads or os.cpu_count() or o.num_threads()   # torchrun exports OMP_NUM_THREADS=1: ask explicitly
    rng = np.random.default_rng(42)
    probe = rng.random((cores * 2, D))
    t0 = time.perf_counter()
    o.refpath_eval(OG, probe, cores)
    per_point = (time.perf_counter() - t0) / len(probe)
    n = int(max(cores * 2, min(200000, seconds_target / max(per_point, 1e-9))))
    n -= n % cores
    X = rng.random((n, D))
    times = []
    s = 0
    while s < warmup + steps:
        t0 = time.perf_counter()
        y = o.refpath_eval(OG, X, cores)
        dt = time.perf_counter() - t0
        if dt < 0.6 * seconds_target and s == 0 and n < 200000:
            # the probe over-estimated the per-point cost (cold caches): grow the sample to the target
            n = int(min(200000, n * seconds_target / max(dt, 1e-3)))
            n -= n % cores
            X = rng.random((n, D))
            continue
        if s >= warmup:
            times.append(dt)
        s += 1
    return n / statistics.mean(times), cores, n, statistics.mean(times), y, X


def _claim_stdout():
    """The contract is ONE JSON line on stdout. Libraries (NCCL prints its version banner from C) also
    write to fd 1, so keep a private copy of the real stdout for the JSON line and point fd 1 at stderr."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(real, "w")


def main():
    out_stream = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--points", type=float, default=1e8, help="total query points per step (all ranks)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--profile", action="store_true",
                    help="short run for ncu: no e2e / cpu legs, warm-up as given (numbers printed are not bench values)")
    args = ap.parse_args()
    M_total = int(args.points)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    W = max(args.warmup, 0)
    K = max(args.steps, 1)
    if args.profile:
        args.no_e2e = args.no_cpu = True
    Wd = W if args.profile else max(W, 3)   # timing rule: at least 3 warm-up steps for a reported number

    if args.impl == "reference":
        # the reference's own CPU implementation of the path on this box's host cores. Julia is not in
        # the image, so this is the oracle port of its algorithm (kind = "port"); rank 0 only.
        if rank != 0:
            return
        Gh = build_oracle_grid()
        rate, cores, n, sec, _, _ = cpu_reference_rate(Gh, seconds_target=15.0, steps=K, warmup=min(W, 1))
        line = {
            "impl": "reference", "metric": "interpolant evaluations per second", "value": rate, "unit": "points/s",
            "n_gpus": args.gpus, "steps": K, "warmup": min(W, 1), "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "nodes": int(len(Gh)), "points_per_step": n,
                       "note": "bounded sample of the 1e8-point workload (the reference streams 224 MB per point, "
                               "src/class.jl:638-642: 1e8 points would take ~1e5 s); per-point cost is independent of M, "
                               "so steps / points differ from the product arm by design",
                       "generous_to_reference": "-O2 OpenMP C port of the reference's loop structure; Julia's broadcast "
                                                "additionally allocates N doubles per point"},
            "products_per_s": rate * len(Gh),
            "cpu_baseline": {"value": rate, "unit": "points/s", "cores": cores, "kind": "port",
                             "sample": f"{n} of {M_total} points, restated reference algorithm "
                                       "(src/class.jl:635-642: copy coefs, D passes over Int64 N x D, pairwise sum), "
                                       "OpenMP over points; Julia Threads.nthreads() not measurable (no Julia)"},
            "e2e": {"value": rate, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line), file=out_stream, flush=True)
        return

    import torch
    import torch.distributed as dist

    import asg_b200 as asg

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = pin_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    asg.init(local_rank)
    lib = asg._capi.load()

    # ---- grid: rank 0 enumerates, C1 broadcast over NCCL, every rank uploads its replica ----------
    if rank == 0:
        G = build_grid_arrays()
    else:
        G = asg.SparseGridInterpolant(D)
    if world > 1:
        asg.parallel().broadcast_grid(G, src=0)
    h = G._ensure()
    info = G.info()
    N = info["N"]

    # ---- points: contiguous shard of the M_total points, generated on the device ----------------
    lo, hi = asg.parallel().shard_range(M_total, rank, world) if world > 1 else (0, M_total)
    M = hi - lo
    gen = torch.Generator(device=dev)
    gen.manual_seed(42 + rank)
    X = torch.rand((M, D), dtype=torch.float64, device=dev, generator=gen)
    if os.environ.get("ASG_BENCH_SORT"):  # experiment: points pre-sorted by a coarse cell key (2 bits per dimension)
        bits = int(os.environ["ASG_BENCH_SORT"])
        key = torch.zeros(M, dtype=torch.int64, device=dev)
        for d in range(D):
            key = (key << bits) | (X[:, d] * (1 << bits)).to(torch.int64).clamp_(0, (1 << bits) - 1)
        X = X[torch.argsort(key)].contiguous()
        del key
    # C2: the all-gather of the result shards is part of every step for N > 1, inside the timed region;
    # the walk writes its shard straight into the padded send buffer
    per = -(-M_total // world)
    pad = torch.zeros(per, dtype=torch.float64, device=dev)
    out = pad[:M]
    full = torch.empty(per * world, dtype=torch.float64, device=dev) if world > 1 else None
    stream = torch.cuda.current_stream()

    def step_device():
        asg._capi.check(lib.asg_eval_device(h, C.c_void_p(X.data_ptr()), M, D, 1, -1, C.c_void_p(out.data_ptr()),
                                            C.c_void_p(stream.cuda_stream)))
        if world > 1:
            dist.all_gather_into_tensor(full, pad)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(Wd):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = asg.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ms, call_ms = [], []
    barrier()
    e0.record(stream)
    for _ in range(K):
        step_device()
    e1.record(stream)
    barrier()
    launches = asg.launch_count() - l0
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    # per-launch kernel time of the dominant kernel, measured live with CUDA events on the launch stream
    for _ in range(3):
        step_device()
        ms = C.c_double()
        asg._capi.check(lib.asg_last_walk_ms(C.byref(ms)))     # the walk kernel alone
        kern_ms.append(ms.value)
        asg._capi.check(lib.asg_last_kernel_ms(C.byref(ms)))   # point ordering + walk
        call_ms.append(ms.value)
    # the gather alone (reported; it is already inside ms_per_step)
    gather_ms = None
    if world > 1:
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        g0.record()
        dist.all_gather_into_tensor(full, pad)
        g1.record()
        barrier()
        gather_ms = g0.elapsed_time(g1)
    t = torch.tensor([ms_total, float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_total, launches = float(tmax[0]), int(tsum[1])
    ms_per_step = ms_total / K
    value = M_total / (ms_per_step * 1e-3)
    checksum = float(out.sum().item())
    del full
    # a sample of THE BENCHMARKED LAUNCH itself (points and results of the timed, device-resident, paired walk) for the
    # oracle comparison below
    sidx = torch.randint(0, M, (1000,), device=dev, generator=gen)
    X_sample, y_sample = X[sidx].cpu().numpy(), out[sidx].cpu().numpy()

    # ---- e2e: the public API with pinned host buffers (H2D and D2H inside the timed region) ------
    e2e = None
    if not args.no_e2e:
        Xh = torch.empty((M, D), dtype=torch.float64, pin_memory=True)
        Xh.copy_(X)
        oh = torch.empty(M, dtype=torch.float64, pin_memory=True)
        Xn, on = Xh.numpy(), oh.numpy()
        del X
        torch.cuda.empty_cache()
        G.evaluate(Xn, out=on)  # warm-up (allocates the pipeline buffers)
        ke = min(K, 3)
        barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            G.evaluate(Xn, out=on)
        barrier()
        dt = (time.perf_counter() - t0) / ke
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt[0])
        ok = abs(float(on.sum()) - checksum) <= 1e-9 * max(1.0, abs(checksum))
        e2e = {"value": M_total / dt, "unit": "points/s", "h2d_bytes_per_step": M_total * D * 8,
               "d2h_bytes_per_step": M_total * 8, "ms_per_step": dt * 1e3, "steps": ke,
               "matches_device_run": bool(ok)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_eval_block<2,false,512>; k_eval_stream on other grids) ----
    # The binding resource of the walk is the shared-memory data pipe: one 128-byte WAVEFRONT per clock per SM (ncu:
    # l1tex__data_pipe_lsu_wavefronts_mem_shared). Counts per point (wavefronts, FP64 thread instructions, DRAM bytes)
    # are MEASURED: they come from the committed ncu capture of this same kernel (profiles/*_summary.json, written by
    # profiles/summarize_ncu.py); times, clocks and the FP64 peak are measured live in this run. So
    # frac * peak * kernel_ms == the ncu wavefront count scaled to this launch's points, by construction.
    fma_rate = C.c_double()
    asg._capi.check(lib.asg_measure_fp64_peak(C.byref(fma_rate)))
    block = bool(info.get("block")) and os.environ.get("ASG_EVAL_KERNEL", "b")[0] == "b"
    kname = "k_eval_block<2,false,512,PAIR>" if block else f"k_eval_stream<{D},R,false>"
    prof_names = (["r02z_eval_block_d10_summary.json", "r02_eval_block_d10_summary.json", "r01b_final_eval_block_d10_summary.json"] if block
                  else ["r01_final_eval_stream_d10_summary.json"])
    k_ms = statistics.median(kern_ms)
    pts_per_launch = M
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    alg_bytes = pts_per_launch * (8 * D + 8)
    props = torch.cuda.get_device_properties(dev)
    n_sm = props.multi_processor_count
    sm_hz = (clocks["sm_mhz"] if clocks and clocks.get("sm_mhz") else peaks.get("sm_max_mhz", 1965.0)) * 1e6
    instr_model, flops_model, leaves = (fp64_op_counts_block(G.levels, int(info["block_records"])) if block
                                        else fp64_op_counts(G.levels))
    roofline = {
        "bound": "smem",  # shared-memory wavefronts; neither "hbm" nor "tensor" describes a sparse gather-product
        "kernel": kname, "kernel_ms": k_ms, "points_per_launch": pts_per_launch,
        "point_order_ms": statistics.median(call_ms) - k_ms,
        "achieved": None, "peak": n_sm * sm_hz * 128.0 / 1e9, "unit": "GB/s", "frac": None, "traffic": None,
        "peak_source": f"{n_sm} SMs x 128 B/clk x SM clock sampled under load in this run ({sm_hz / 1e6:.0f} MHz)",
        "terms_per_point": leaves,
        "hbm": {"achieved": alg_bytes / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes_per_point": 8 * D + 8,
                "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"},
        "dense_equivalent": {"products_per_s": N * pts_per_launch / (k_ms * 1e-3),
                             "fp64_instr_per_product": 3 * D + 1,
                             "frac_of_fp64_peak": (N * (3 * D + 1) * pts_per_launch / (k_ms * 1e-3)) / fma_rate.value,
                             "note": "N*M/t with every (node, point) pair counted, SURVEY.md 8d; > 1 because "
                                     "only the one supporting node per subspace is visited; not a roofline"},
    }
    for prof_name in prof_names:
        prof = os.path.join(ROOT, "profiles", prof_name)
        if not os.path.exists(prof):
            continue
        try:
            pj = json.load(open(prof))
            k0 = pj["kernels"][0]
            ppl = float(pj["points_per_launch"])
            roofline["counts_source"] = f"profiles/{prof_name} (ncu --set full of this kernel, {int(ppl)} points per launch)"
            roofline["ncu_kernel"] = k0.get("kernel")
            # ncu capture may be of a launch with another point count: per-point counts scale, DRAM bytes too
            roofline["traffic"] = pj["dram_bytes_per_launch"] * pts_per_launch / ppl
            if "smem_wavefronts" in k0:
                wpp = k0["smem_wavefronts"] / ppl
                roofline["smem_wavefronts_per_point"] = wpp
                roofline["smem_ld_instructions_per_32_points"] = k0.get("smem_ld_warp_instructions", 0.0) / ppl * 32.0
                roofline["achieved"] = wpp * pts_per_launch * 128.0 / (k_ms * 1e-3) / 1e9
                roofline["frac"] = roofline["achieved"] / roofline["peak"]
                roofline["ncu_frac_same_counter"] = k0.get("smem_wavefront_frac")
            if "fp64_thread_inst" in k0:
                fpp = k0["fp64_thread_inst"] / ppl
                roofline["fp64"] = {"thread_instructions_per_point": fpp,
                                    "achieved": fpp * pts_per_launch / (k_ms * 1e-3) / 1e12, "unit": "Tinstr/s",
                                    "peak": fma_rate.value / 1e12,
                                    "frac": fpp * pts_per_launch / (k_ms * 1e-3) / fma_rate.value,
                                    "peak_source": "DFMA microbenchmark measured live in this run (64 lanes/clk/SM)",
                                    "ncu_pipe_fp64_pct": k0.get("pipe_fp64_pct"),
                                    "model_instr_per_point": instr_model,
                                    "model_note": "analytic count of the FP64 operations in the source (stages, hats, "
                                                  "Horner steps, terms); ptxas executes fewer: level-1 terms of a pole "
                                                  "are plain loads and common sub-expressions of the two points' hats "
                                                  "are shared -- the measured count is the one used"}
            pipes = {"issue_slots": k0.get("issue_slots_busy_pct"), "smem_lsu_wavefronts": k0.get("smem_wavefronts_pct"),
                     "alu": k0.get("pipe_alu_pct"), "fma_int": k0.get("pipe_fma_pct"), "fp64": k0.get("pipe_fp64_pct"),
                     "lsu": k0.get("pipe_lsu_pct"), "dram": k0.get("dram_throughput_pct")}
            pipes = {k: v for k, v in pipes.items() if v is not None}
            roofline["ncu_pipe_utilisation_pct"] = pipes
            roofline["ncu_binding_pipe"] = max(pipes, key=pipes.get) if pipes else None
            break
        except Exception as ex:  # a malformed summary must not kill the bench line
            roofline["counts_error"] = repr(ex)
    if roofline["frac"] is None:  # no capture of this kernel in profiles/: fall back to the HBM figure
        roofline.update({"bound": "hbm", "achieved": roofline["hbm"]["achieved"], "peak": hbm_peak,
                         "frac": roofline["hbm"]["frac"], "note": "no ncu capture for this kernel: algorithmic HBM bytes"})

    cpu = None
    if world == 1 and not args.no_cpu:
        rate, cores, n, sec, ycpu, Xcpu = cpu_reference_rate(G, seconds_target=15.0)
        ygpu = G.evaluate(Xcpu)
        # parity of the benchmarked configuration against the long-double oracle (src/class.jl:629-643), ASSERTED with
        # the tolerance of the north star: 1e-12 relative or 1e-14 * max(1, sum |c phi|) absolute
        from oracle import oracle as o
        OG = o.OracleGrid(D)
        OG.levels, OG.indices, OG.depths, OG.coefs, OG.fvals = G.levels, G.indices, G.depths, G.coefs, G.fvals
        ns = min(len(Xcpu), 1000)
        want, sc = OG.evaluate(Xcpu[:ns], return_abssum=True)
        err = np.abs(ygpu[:ns] - want)
        okp = (err <= 1e-12 * np.abs(want)) | (err <= 1e-14 * np.maximum(1.0, sc))
        if not np.all(okp):
            raise SystemExit(f"PARITY FAILURE on the benchmarked grid: {np.count_nonzero(~okp)} of {ns} points out of "
                             f"tolerance, worst error {float(err.max()):.3e}")
        # ... and of the timed launch itself: 1 000 points drawn from the device-resident batch with the results the
        # benchmarked (ordered, paired) walk wrote for them
        want2, sc2 = OG.evaluate(X_sample, return_abssum=True)
        err2 = np.abs(y_sample - want2)
        ok2 = (err2 <= 1e-12 * np.abs(want2)) | (err2 <= 1e-14 * np.maximum(1.0, sc2))
        if not np.all(ok2):
            raise SystemExit(f"PARITY FAILURE in the benchmarked launch: {np.count_nonzero(~ok2)} of {len(want2)} sampled points "
                             f"out of tolerance, worst error {float(err2.max()):.3e}")
        cpu = {"value": rate, "unit": "points/s", "cores": cores, "kind": "port",
               "sample": f"{n} of {M_total} points in {sec:.1f} s, restated reference algorithm "
                         "(src/class.jl:635-642), -O2 OpenMP over points (generous to the reference: Julia's broadcast "
                         "also allocates N doubles per point)",
               "max_abs_diff_vs_gpu": float(np.max(np.abs(ygpu - ycpu))),
               "parity_vs_long_double_oracle": {"points": int(ns), "max_abs_err": float(err.max()),
                                                "timed_launch_sample_points": int(len(want2)),
                                                "timed_launch_sample_max_abs_err": float(err2.max()),
                                                "tolerance": "1e-12 rel | 1e-14 * max(1, sum|c*phi|) abs", "asserted": True}}

    # a second CPU row from the reference ITSELF, if somebody has run tests/golden/dump_reference.jl (Julia >= 1.9): its
    # header carries Threads.nthreads() and the wall time of G.(points) on cfg1 (737 nodes) -- another grid than the
    # headline's, so it is reported next to cpu_baseline, not instead of it
    if cpu is not None:
        dump = os.path.join(ROOT, "tests", "golden", "julia_cfg1.bin")
        if os.path.exists(dump):
            import struct
            with open(dump, "rb") as fh:
                hd = fh.read(56)
            Dj, Nj, Mj, nth = struct.unpack_from("<4q", hd, 0)
            wt, wa, we = struct.unpack_from("<3d", hd, 32)
            cpu["julia_reference_cfg1"] = {"value": Mj / we if we > 0 else None, "unit": "points/s", "threads": int(nth),
                                           "nodes": int(Nj), "points": int(Mj), "train_s": wt, "kind": "reference",
                                           "note": "the unmodified reference's G.(points) on cfg1 (D = 3, 737 nodes)"}

    line = {
        "metric": "interpolant evaluations per second", "value": value, "unit": "points/s", "n_gpus": world,
        "steps": K, "warmup": Wd, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "nodes": int(N), "subspaces": int(info["subspaces"]),
                   "points_per_step": M_total, "points_per_gpu": M, "parallelism": f"grid replicated, points sharded x{world}",
                   "host_cores_bound_to_gpu_numa_node": numa,
                   "l2": "inputs larger than L2 (80 B/point * points_per_gpu >> 126 MB), no flush needed",
                   "engine": ("subspace/block" if block else "subspace/stream") if info["path"] == 2 else "sweep",
                   "point_order": "cell-key sorted inside the timed region (asg_sort.cu)" if block else "as given",
                   "timed_region": "point ordering + walk" + (" + all-gather of the result shards (C2)" if world > 1 else "")},
        "products_per_s": value * N,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "e2e": e2e,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "gather_ms": gather_ms,
        "checksum": checksum,
    }
    print(json.dumps(line), file=out_stream, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
