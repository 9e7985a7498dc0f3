#!/usr/bin/env python
"""Pinned host <-> device copy bandwidth of this box: every visible GPU alone, then all of them at once.

The end-to-end benchmark (bench.py: e2e) moves 80 B in and 8 B out per point through PCIe; at 8 GPUs the eight
H2D streams share the host's memory system and PCIe fabric. This script measures that ceiling directly (one
cudaMemcpyAsync per 256 MiB chunk, one stream and one pinned buffer per GPU) so that the e2e scaling can be read
against it:   python benchmarks/pcie_bw.py [--mib 1024] [--reps 5]  ->  one JSON line
"""
import argparse
import json
import threading
import time

import torch


def copy_loop(dev, host, devbuf, reps, direction, chunk, out, barrier):
    torch.cuda.set_device(dev)
    st = torch.cuda.Stream(device=dev)
    n = host.numel()
    with torch.cuda.stream(st):
        for _ in range(2):  # warm-up
            devbuf.copy_(host, non_blocking=True) if direction == "h2d" else host.copy_(devbuf, non_blocking=True)
        st.synchronize()
        if barrier is not None:
            barrier.wait()
        t0 = time.perf_counter()
        for _ in range(reps):
            for o in range(0, n, chunk):
                if direction == "h2d":
                    devbuf[o:o + chunk].copy_(host[o:o + chunk], non_blocking=True)
                else:
                    host[o:o + chunk].copy_(devbuf[o:o + chunk], non_blocking=True)
        st.synchronize()
        dt = time.perf_counter() - t0
    out[dev] = n * host.element_size() * reps / dt / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    ng = torch.cuda.device_count()
    n = a.mib * (1 << 20) // 8
    chunk = (256 << 20) // 8
    hosts = [torch.empty(n, dtype=torch.float64, pin_memory=True).fill_(1.0) for _ in range(ng)]
    devs = [torch.empty(n, dtype=torch.float64, device=f"cuda:{d}") for d in range(ng)]
    res = {"gpus": ng, "mib_per_copy": a.mib, "chunk_mib": 256, "reps": a.reps}
    for direction in ("h2d", "d2h"):
        alone = {}
        for d in range(ng):
            copy_loop(d, hosts[d], devs[d], a.reps, direction, chunk, alone, None)
        res[direction + "_alone_gbs"] = [round(alone[d], 2) for d in range(ng)]
        if ng > 1:
            together = {}
            bar = threading.Barrier(ng)
            th = [threading.Thread(target=copy_loop, args=(d, hosts[d], devs[d], a.reps, direction, chunk, together, bar))
                  for d in range(ng)]
            [t.start() for t in th]
            [t.join() for t in th]
            res[direction + "_concurrent_gbs"] = [round(together[d], 2) for d in range(ng)]
            res[direction + "_concurrent_total_gbs"] = round(sum(together.values()), 2)
    # both directions at once on every GPU (what the evaluation pipeline does)
    if ng >= 1:
        both = {}
        outs = [dict(), dict()]
        bar = threading.Barrier(2 * ng)
        hosts2 = [torch.empty(n // 8, dtype=torch.float64, pin_memory=True) for _ in range(ng)]
        devs2 = [torch.empty(n // 8, dtype=torch.float64, device=f"cuda:{d}") for d in range(ng)]
        th = [threading.Thread(target=copy_loop, args=(d, hosts[d], devs[d], a.reps, "h2d", chunk, outs[0], bar)) for d in range(ng)]
        th += [threading.Thread(target=copy_loop, args=(d, hosts2[d], devs2[d], a.reps * 8, "d2h", chunk, outs[1], bar)) for d in range(ng)]
        [t.start() for t in th]
        [t.join() for t in th]
        res["duplex_h2d_total_gbs"] = round(sum(outs[0].values()), 2)
        res["duplex_d2h_total_gbs"] = round(sum(outs[1].values()), 2)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
