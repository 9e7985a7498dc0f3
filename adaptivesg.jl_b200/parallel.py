"""Multi-GPU layer: the grid is REPLICATED, query points are SHARDED.

One process per GPU (``torch.distributed``; NCCL over NVLink on the GPU box, gloo in CPU tests).
The evaluation itself has no data-path collective: points are independent. Collectives appear only
  C1  ``broadcast_grid``  -- node arrays root -> all, once per grid version,
  C2  ``gather_results``  -- the M x 8 byte result shards back to every rank / the root.
``train``/``adapt`` levels use the same pattern: shard the candidates, all-gather alpha, every rank
appends the same rows.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_range(M: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous ceil(M / world) ranges: preserves the output order under concatenation."""
    per = -(-M // world)
    lo = min(M, rank * per)
    return lo, min(M, lo + per)


def _dev():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def broadcast_grid(G, src: int = 0, packed: bool | None = None):
    """C1: make every rank's SparseGridInterpolant identical to rank `src`'s (arrays + bounds).

    With ``packed`` (default: on whenever the backend is NCCL, i.e. every rank has its GPU) the source also ships the
    device-format blob of its uploaded grid (``asg_grid_export``): the other ranks install it with
    ``asg_grid_import_checked`` instead of validating, sorting and packing the same node arrays once more per rank
    (0.13 s instead of 0.3 s each for the 652 065-node benchmark grid)."""
    import ctypes as C

    from . import _capi
    from .interpolant import serialize

    dev = _dev()
    D = G.ndims
    is_src = dist.get_rank() == src
    if packed is None:
        packed = dist.get_backend() == "nccl"
    blob = serialize(G, packed=True)["packed"] if (packed and is_src and len(G.levels)) else None
    meta = torch.tensor([len(G.levels), 0 if blob is None else blob.size] if is_src else [0, 0], dtype=torch.int64, device=dev)
    dist.broadcast(meta, src)
    N, nblob = int(meta[0].item()), int(meta[1].item())

    def bc(arr, shape, dtype):
        t = torch.from_numpy(np.ascontiguousarray(arr, dtype=dtype)).to(dev) if is_src else \
            torch.empty(shape, dtype=torch.from_numpy(np.empty(0, dtype=dtype)).dtype, device=dev)
        dist.broadcast(t, src)
        return t.cpu().numpy()

    bounds = bc(np.concatenate([G.lb, G.ub]), (2 * D,), np.float64)
    levels = bc(G.levels, (N, D), np.int64)
    indices = bc(G.indices, (N, D), np.int64)
    depths = bc(G.depths, (N,), np.int64)
    fvals = bc(G.fvals, (N,), np.float64)
    coefs = bc(G.coefs, (N,), np.float64)
    blob_all = bc(blob, (nblob,), np.uint8) if nblob else None
    if not is_src:
        G.lb, G.ub = tuple(bounds[:D].tolist()), tuple(bounds[D:].tolist())
        G.levels, G.indices, G.depths, G.fvals, G.coefs = levels, indices, depths, fvals, coefs
        if blob_all is not None:
            buf = np.ascontiguousarray(blob_all)
            rc = _capi.load().asg_grid_import_checked(
                G._grid(), buf.ctypes.data_as(C.c_void_p), buf.size, N, levels.ctypes.data_as(_capi.i64p),
                indices.ctypes.data_as(_capi.i64p), D, 1, coefs.ctypes.data_as(_capi.f64p))
            if rc == 0:
                object.__setattr__(G, "_dirty", False)
            elif rc not in (_capi.ASG_EUNSUPPORTED, _capi.ASG_ESTATE):
                _capi.check(rc)
    return G


def gather_results(local: np.ndarray, M: int) -> np.ndarray:
    """C2: concatenate the per-rank result shards (shard_range order) on every rank."""
    world = dist.get_world_size()
    dev = _dev()
    per = -(-M // world) if M else 0
    buf = torch.zeros(max(per, 1), dtype=torch.float64, device=dev)
    if len(local):
        buf[: len(local)] = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64)).to(dev)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    sizes = [shard_range(M, r, world) for r in range(world)]
    return np.concatenate([parts[r].cpu().numpy()[: hi - lo] for r, (lo, hi) in enumerate(sizes)])


def evaluate_sharded(G, X: np.ndarray, max_depth: int | None = None, evaluator=None) -> np.ndarray:
    """``[G(x) for x in eachrow(X)]`` with the rows split over the ranks. Every rank passes the same
    X (or at least its own rows); every rank gets the full result. `evaluator` defaults to
    ``G.evaluate`` (the CUDA path); CPU tests inject a stand-in to exercise the sharding logic."""
    M = X.shape[0]
    lo, hi = shard_range(M, dist.get_rank(), dist.get_world_size())
    ev = evaluator if evaluator is not None else (lambda Z: G.evaluate(Z, max_depth=max_depth))
    local = ev(X[lo:hi]) if hi > lo else np.empty(0)
    return gather_results(local, M)
