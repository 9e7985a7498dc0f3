"""Import shim: the package directory is ``adaptivesg.jl_b200`` (a dot is not importable), so this
module loads it under the name ``asg_b200`` and re-exports its namespace."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "adaptivesg.jl_b200")
_NAME = "_asg_b200_pkg"

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_PKG_DIR, "__init__.py"),
                                                   submodule_search_locations=[_PKG_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)
_pkg = sys.modules[_NAME]

globals().update({k: getattr(_pkg, k) for k in _pkg.__all__})
_capi = _pkg._capi
interpolant = sys.modules[_NAME + ".interpolant"]
PKG_DIR = _PKG_DIR


def parallel():
    """Lazy import of the torch.distributed helpers (keeps `import asg_b200` torch-free)."""
    import importlib
    return importlib.import_module(_NAME + ".parallel")
