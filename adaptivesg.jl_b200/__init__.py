"""asg-b200: B200-native engine for AdaptiveSG.jl's hot path (batched hat-basis sparse-grid
evaluation). The directory name carries a dot, so import it through the repo-root shim::

    import asg_b200 as asg
    G = asg.SparseGridInterpolant(3)
    asg.train(G, f, 7, 5)
    y = G.evaluate(X)

Layout: ``csrc/`` CUDA kernels + C ABI (-> ``lib/libasg_cuda.so``), ``_capi.py`` ctypes binding,
``interpolant.py`` host mirror of the reference interface, ``parallel.py`` point sharding over
``torch.distributed``, ``julia/`` the ccall shim for the reference package.
"""
from . import _capi
from ._capi import AsgError, InvalidNodeError, LIB_PATH, context_devices, init, init_devices, launch_count, shutdown
from .interpolant import (SparseGridInterpolant, adapt, allindices_of_level, allnodes1d, basis, boundaryflag, child,
                          coefficients, dehierarchization_matrix, deserialize, extrapolate, fraction2li, get_x,
                          ghost_distance, ghost_neighbor, hierarchization_matrix, hierarchize, isbig2add, isboundary,
                          isvalid, maxlevels, nodedepth, parent, perturb, phi, rsg, scale, serialize, stack, train)

__all__ = [
    "SparseGridInterpolant", "rsg", "train", "adapt", "coefficients", "basis", "dehierarchization_matrix",
    "hierarchization_matrix", "serialize", "deserialize", "maxlevels", "stack", "nodedepth", "isvalid",
    "allindices_of_level", "child", "isbig2add", "hierarchize", "extrapolate", "AsgError", "InvalidNodeError", "init", "launch_count",
    "LIB_PATH", "phi", "scale", "get_x", "parent", "perturb", "isboundary", "boundaryflag", "allnodes1d", "fraction2li",
    "ghost_neighbor", "ghost_distance", "init_devices", "context_devices", "shutdown",
]
