"""CPU test of the N > 1 path: world_size-2 gloo processes exercise the sharding, the grid
broadcast (C1) and the result gather (C2). The evaluator is injected (a cheap deterministic
function of the rows) because no GPU is present here; the CUDA evaluator is covered by -m gpu."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import asg_b200 as asg
    par = asg.parallel()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        G = asg.SparseGridInterpolant(3, lb=[0, 1, 2], ub=[1, 3, 5]) if rank == 0 else asg.SparseGridInterpolant(3)
        if rank == 0:
            asg.rsg(G, 5, 4)
            G.coefs = np.arange(len(G), dtype=np.float64)
            G.fvals = -G.coefs
        par.broadcast_grid(G, src=0)
        ok = (len(G) == 111 or len(G) > 0) and G.lb == (0.0, 1.0, 2.0) and G.ub == (1.0, 3.0, 5.0)
        ok &= bool(np.array_equal(G.coefs, np.arange(len(G)))) and G.levels.shape == (len(G), 3)
        for M in (0, 1, 7, 1000, 1001):
            X = np.random.default_rng(M).random((M, 3))
            want = X @ np.array([1.0, -2.0, 0.5]) + G.coefs.sum()
            got = par.evaluate_sharded(G, X, evaluator=lambda Z: Z @ np.array([1.0, -2.0, 0.5]) + G.coefs.sum())
            ok &= got.shape == (M,) and bool(np.array_equal(got, want))
            lo, hi = par.shard_range(M, rank, world)
            ok &= 0 <= lo <= hi <= M
        q.put((rank, bool(ok), len(G)))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1] and all(r[1] for r in res)
    assert res[0][2] == res[1][2] > 0


def test_shard_range_covers_everything():
    sys.path.insert(0, ROOT)
    import asg_b200 as asg
    par = asg.parallel()
    for M in (0, 1, 5, 8, 100000001):
        for world in (1, 2, 4, 8):
            r = [par.shard_range(M, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == M
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
