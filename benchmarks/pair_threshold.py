#!/usr/bin/env python
"""Where does PAIR mode of the BLOCK walk start to pay? cfg3 grid, device-resident batches of several sizes, the
same batch walked with ASG_BLOCK_PAIR=0 (unpaired) and forced on; prints ms per call (point ordering + walk +
follow-up) and points/s. One JSON line per batch size."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import asg_b200 as asg  # noqa: E402
import bench  # noqa: E402


def main():
    asg.init(0)
    lib = asg._capi.load()
    G = bench.build_grid_arrays()
    h = G._ensure()
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(7)
    sizes = [int(float(a)) for a in sys.argv[1:]] or [1 << 20, 1 << 21, 1 << 22, 1 << 23, 12500000, 1 << 24, 1 << 25, 100000000]
    for M in sizes:
        X = torch.rand((M, bench.D), dtype=torch.float64, device=dev, generator=gen)
        out = torch.empty(M, dtype=torch.float64, device=dev)
        st = torch.cuda.current_stream()
        res = {"points": M, "points_per_cell": M / 4.0 ** bench.D}
        ref = None
        for mode, env in (("unpaired", "0"), ("paired", "1e-9")):
            os.environ["ASG_BLOCK_PAIR"] = env
            ms = []
            for rep in range(4):
                asg._capi.check(lib.asg_eval_device(h, C.c_void_p(X.data_ptr()), M, bench.D, 1, -1, C.c_void_p(out.data_ptr()),
                                                    C.c_void_p(st.cuda_stream)))
                v = C.c_double()
                asg._capi.check(lib.asg_last_kernel_ms(C.byref(v)))
                if rep:
                    ms.append(v.value)
            res[mode + "_ms"] = float(np.median(ms))
            res[mode + "_points_per_s"] = M / (res[mode + "_ms"] * 1e-3)
            y = out.clone()
            if ref is None:
                ref = y
            else:
                res["bit_identical"] = bool(torch.equal(ref, y))
        res["speedup"] = res["unpaired_ms"] / res["paired_ms"]
        print(json.dumps(res), flush=True)
        del X, out


if __name__ == "__main__":
    main()
