"""ctypes binding of libasg_cuda.so (include/asg_cuda.h) -- the same calls a Julia `ccall` makes.

There is no CPU fallback: if the shared library is missing or no sm_100 device is present every
compute call raises ``AsgError``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# ASG_CUDA_LIB: another build of the library (A/B measurements of kernel variants on one machine)
LIB_PATH = os.environ.get("ASG_CUDA_LIB") or os.path.join(_HERE, "lib", "libasg_cuda.so")

ASG_OK = 0
ASG_EINVAL = -1
ASG_EINVALID_NODE = -2
ASG_ENODEVICE = -3
ASG_ECUDA = -4
ASG_ENOMEM = -5
ASG_EUNSUPPORTED = -6
ASG_ESTATE = -7
ASG_PATH_AUTO, ASG_PATH_SWEEP, ASG_PATH_SUBSPACE, ASG_PATH_TRIE, ASG_PATH_STREAM, ASG_PATH_BLOCK = 0, 1, 2, 3, 4, 5

_CODE_NAMES = {
    ASG_EINVAL: "ASG_EINVAL", ASG_EINVALID_NODE: "ASG_EINVALID_NODE", ASG_ENODEVICE: "ASG_ENODEVICE",
    ASG_ECUDA: "ASG_ECUDA", ASG_ENOMEM: "ASG_ENOMEM", ASG_EUNSUPPORTED: "ASG_EUNSUPPORTED",
    ASG_ESTATE: "ASG_ESTATE",
}


class AsgError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{_CODE_NAMES.get(code, code)}: {msg}")
        self.code = code


class InvalidNodeError(AsgError, ValueError):
    """The reference throws ArgumentError("invalid level-index pair") here (src/math.jl:355, 474)."""


class GridInfo(C.Structure):
    _fields_ = [
        ("D", C.c_int32), ("canonical", C.c_int32), ("path", C.c_int32), ("max_level", C.c_int32),
        ("N", C.c_int64), ("max_depth", C.c_int64), ("subspaces", C.c_int64), ("dense_subspaces", C.c_int64),
        ("trie_nodes", C.c_int64), ("coef_slots", C.c_int64), ("hash_slots", C.c_int64),
        ("device_bytes", C.c_int64), ("generation", C.c_int64), ("regular", C.c_int64),
        ("block", C.c_int64), ("block_slots", C.c_int64), ("block_records", C.c_int64),
        ("block_tiles", C.c_int64), ("block_stack", C.c_int64), ("block_pair", C.c_int64),
        ("packs", C.c_int64), ("delta_nodes", C.c_int64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


i64p = C.POINTER(C.c_int64)
f64p = C.POINTER(C.c_double)
vp = C.c_void_p

# name -> (argtypes); every function returns int32 except the two string getters
SIGNATURES = {
    "asg_init": [C.c_int32],
    "asg_init_devices": [C.POINTER(C.c_int32), C.c_int32],
    "asg_context_devices": [C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32)],
    "asg_phi": [C.c_int64, C.c_int32, f64p, i64p, i64p, f64p],
    "asg_shutdown": [],
    "asg_device_count": [C.POINTER(C.c_int32)],
    "asg_launch_count": [i64p],
    "asg_host_register": [vp, C.c_size_t],
    "asg_host_unregister": [vp],
    "asg_grid_create": [C.c_int32, f64p, f64p, C.POINTER(vp)],
    "asg_grid_destroy": [vp],
    "asg_grid_upload": [vp, C.c_int64, i64p, i64p, C.c_int64, C.c_int64, f64p],
    "asg_grid_append": [vp, C.c_int64, i64p, i64p, C.c_int64, C.c_int64, f64p],
    "asg_grid_set_coefs": [vp, C.c_int64, C.c_int64, f64p],
    "asg_grid_scatter_coefs": [vp, C.c_int64, i64p, f64p],
    "asg_grid_get_coefs": [vp, C.c_int64, C.c_int64, f64p],
    "asg_grid_info": [vp, C.POINTER(GridInfo)],
    "asg_grid_set_path": [vp, C.c_int32],
    "asg_eval": [vp, vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, vp],
    "asg_surplus": [vp, vp, C.c_int64, C.c_int64, C.c_int64, vp, C.c_int64, vp],
    "asg_surplus_nodes": [vp, C.c_int64, i64p, f64p, C.c_int64, C.c_int32, f64p],
    "asg_eval_device": [vp, vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, vp, vp],
    "asg_surplus_device": [vp, vp, C.c_int64, C.c_int64, C.c_int64, vp, C.c_int64, vp, vp],
    "asg_eval_async": [vp, vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, vp, C.POINTER(vp)],
    "asg_wait": [vp],
    "asg_nodes_x": [vp, C.c_int64, C.c_int64, f64p, C.c_int64, C.c_int64],
    "asg_nodes_x_li": [vp, C.c_int64, i64p, i64p, C.c_int64, C.c_int64, f64p, C.c_int64, C.c_int64],
    "asg_basis_count": [vp, vp, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_int32, i64p],
    "asg_basis_fill": [vp, vp, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_int32, i64p, i64p, f64p],
    "asg_basis_fill_csc": [vp, vp, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_int32, i64p, i64p, i64p, f64p],
    "asg_basis_fill_csc32": [vp, vp, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_int32, i64p, C.POINTER(C.c_int32),
                             C.POINTER(C.c_int32), f64p],
    "asg_basis_dense": [vp, vp, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_int32, vp, C.c_int64, C.c_int64],
    "asg_basis_dense_device": [vp, vp, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_int32, vp, C.c_int64,
                               C.c_int64, vp],
    "asg_pack_info": [C.c_int32, C.c_int64, i64p, i64p, C.c_int64, C.c_int64, C.POINTER(GridInfo)],
    "asg_measure_fp64_peak": [f64p],
    "asg_grid_export": [vp, vp, C.c_int64, i64p],
    "asg_grid_import": [vp, vp, C.c_int64],
    "asg_grid_import_checked": [vp, vp, C.c_int64, C.c_int64, i64p, i64p, C.c_int64, C.c_int64, f64p],
    "asg_rsg_count": [C.c_int32, C.c_int64, i64p, i64p],
    "asg_rsg_fill": [C.c_int32, C.c_int64, i64p, C.c_int64, i64p, i64p, C.c_int64, C.c_int64],
    "asg_children": [vp, C.c_int64, i64p, i64p, i64p],
    "asg_children_fetch": [vp, C.c_int64, i64p, i64p, C.c_int64, C.c_int64, f64p, C.c_int64, C.c_int64],
    "asg_last_kernel_ms": [f64p],
    "asg_last_walk_ms": [f64p],
}
STRING_GETTERS = ("asg_last_error", "asg_version")

_lib = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """dlopen the in-tree library and declare every prototype of include/asg_cuda.h."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise AsgError(ASG_ENODEVICE, f"{LIB_PATH} is not built (run __graft_entry__.build()); "
                                              "there is no CPU fallback")
            L = C.CDLL(LIB_PATH)
            for name, args in SIGNATURES.items():
                fn = getattr(L, name)
                fn.argtypes = args
                fn.restype = C.c_int32
            for name in STRING_GETTERS:
                getattr(L, name).restype = C.c_char_p
                getattr(L, name).argtypes = []
            _lib = L
    return _lib


def check(rc: int):
    if rc != ASG_OK:
        msg = load().asg_last_error().decode("utf-8", "replace")
        if rc == ASG_EINVALID_NODE:
            raise InvalidNodeError(rc, msg)
        raise AsgError(rc, msg)


_init_device = None
_init_devices = None


def init(device: int | None = None) -> int:
    """asg_init on `device` (default: LOCAL_RANK or 0). Raises AsgError(ASG_ENODEVICE) without a B200."""
    global _init_device, _init_devices
    if device is None:
        if _init_device is not None:
            return _init_device
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if _init_devices is not None and len(_init_devices) > 1 and _init_devices[0] == int(device):
        return _init_device  # a wider context with this primary device is already up
    check(load().asg_init(int(device)))
    _init_device = int(device)
    _init_devices = [int(device)]
    return _init_device


def init_devices(devices) -> list:
    """asg_init_devices: one process drives several GPUs of the box; every grid keeps a replica per device and the
    host-buffer calls (evaluate / surplus / train) split their points over them. devices[0] is the primary."""
    global _init_device, _init_devices
    devs = [int(d) for d in devices]
    arr = (C.c_int32 * len(devs))(*devs)
    check(load().asg_init_devices(arr, len(devs)))
    _init_device, _init_devices = devs[0], devs
    return devs


def context_devices() -> list:
    n = C.c_int32()
    arr = (C.c_int32 * 16)()
    check(load().asg_context_devices(arr, 16, C.byref(n)))
    return [int(arr[k]) for k in range(n.value)]


def shutdown():
    global _init_device, _init_devices
    check(load().asg_shutdown())
    _init_device = _init_devices = None


def device_ready() -> bool:
    """True once asg_init has succeeded in this process."""
    return _init_device is not None


def launch_count() -> int:
    n = C.c_int64()
    check(load().asg_launch_count(C.byref(n)))
    return n.value
