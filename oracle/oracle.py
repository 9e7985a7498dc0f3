"""CPU oracle for the AdaptiveSG.jl hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this module.  The product (``adaptivesg.jl_b200``) never does.

PARITY STATUS: "parity unpinned" by reference-produced artefacts (the reference has no tests or
fixtures and Julia cannot run in this image).  See the header of ``asg_oracle.c`` for what pins it.

This module is a thin ctypes layer over ``libasg_oracle.so`` (the C restatement) plus a Python
restatement of the host-side training logic that surrounds the hot path:

* ``OracleGrid``            -- src/class.jl:281-319 (the node store)
* ``OracleGrid.rsg``        -- src/train.jl:40-78
* ``OracleGrid.train``      -- src/train.jl:112-195
* ``OracleGrid.adapt``      -- src/train.jl:235-463
* ``OracleGrid.__call__``   -- src/class.jl:747-761 -> 629-643
* ``OracleGrid.basis``      -- src/class.jl:434-538 (documented intent, SURVEY.md defect B1)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libasg_oracle.so")

c_i64p = C.POINTER(C.c_int64)
c_f64p = C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "asg_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libasg_oracle.so"])
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.asgo_power2.restype = C.c_int64
        L.asgo_power2.argtypes = [C.c_int64]
        L.asgo_phi1.restype = C.c_double
        L.asgo_phi1.argtypes = [C.c_double]
        L.asgo_phi.restype = C.c_int
        L.asgo_phi.argtypes = [C.c_double, C.c_int64, C.c_int64, c_f64p]
        L.asgo_get_x1.restype = C.c_int
        L.asgo_get_x1.argtypes = [C.c_int64, C.c_int64, c_f64p]
        L.asgo_isvalid.restype = C.c_int
        L.asgo_isvalid.argtypes = [C.c_int, c_i64p, c_i64p, C.c_int64]
        L.asgo_child1.restype = C.c_int
        L.asgo_child1.argtypes = [C.c_int64, C.c_int64, C.c_int, c_i64p, c_i64p]
        L.asgo_parent1.restype = C.c_int
        L.asgo_parent1.argtypes = [C.c_int64, C.c_int64, c_i64p, c_i64p]
        L.asgo_allindices_of_level.restype = C.c_int64
        L.asgo_allindices_of_level.argtypes = [C.c_int64, c_i64p]
        L.asgo_fraction2li.restype = C.c_int
        L.asgo_fraction2li.argtypes = [C.c_int64, C.c_int64, c_i64p, c_i64p]
        L.asgo_isbig2add.restype = C.c_int
        L.asgo_isbig2add.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
        L.asgo_stack.restype = C.c_int
        L.asgo_stack.argtypes = [C.c_int, C.c_int64, c_i64p, c_i64p, C.c_int64, C.c_int64, c_f64p, c_f64p, c_f64p]
        L.asgo_interpolate_batch.restype = C.c_int
        L.asgo_interpolate_batch.argtypes = [
            C.c_int, C.c_int64, c_i64p, c_i64p, C.c_int64, C.c_int64, c_f64p, c_f64p, c_f64p,
            c_f64p, C.c_int64, C.c_int64, C.c_int64, c_i64p, C.c_int64, c_f64p, c_f64p]
        L.asgo_basis_dense.restype = C.c_int
        L.asgo_basis_dense.argtypes = [
            C.c_int, C.c_int64, c_i64p, c_i64p, C.c_int64, C.c_int64, c_f64p, c_f64p,
            c_f64p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_double, c_f64p]
        L.asgo_rsg.restype = C.c_int64
        L.asgo_rsg.argtypes = [C.c_int, C.c_int64, c_i64p, c_i64p, c_i64p, c_i64p]
        L.asgo_refpath_batch.restype = C.c_int
        L.asgo_refpath_batch.argtypes = [
            C.c_int, C.c_int64, c_i64p, c_i64p, c_f64p, c_f64p, c_f64p,
            c_f64p, C.c_int64, C.c_int64, C.c_int64, c_f64p, C.c_int]
        L.asgo_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _i64(a):
    return a.ctypes.data_as(c_i64p)


def _f64(a):
    return a.ctypes.data_as(c_f64p)


class ArgumentError(ValueError):
    """Stands in for Julia's ArgumentError (src/math.jl:355, 474)."""


# ---------------------------------------------------------------------------------------------
# scalar helpers (thin wrappers; used by the known-answer tests)
# ---------------------------------------------------------------------------------------------
def power2(x: int) -> int:
    return int(lib().asgo_power2(int(x)))


def phi(x: float, l: int | None = None, i: int | None = None) -> float:
    if l is None:
        return float(lib().asgo_phi1(float(x)))
    out = C.c_double()
    if lib().asgo_phi(float(x), int(l), int(i), C.byref(out)) != 0:
        raise ArgumentError(f"invalid level-index pair ({l}, {i})")
    return out.value


def get_x1(l: int, i: int) -> float:
    out = C.c_double()
    if lib().asgo_get_x1(int(l), int(i), C.byref(out)) != 0:
        raise ArgumentError(f"invalid level-index pair ({l}, {i})")
    return out.value


def isvalid(Ls, Is) -> bool:
    Ls = np.ascontiguousarray(Ls, dtype=np.int64)
    Is = np.ascontiguousarray(Is, dtype=np.int64)
    if Ls.shape != Is.shape:
        return False
    return bool(lib().asgo_isvalid(len(Ls), _i64(Ls), _i64(Is), 1))


def child1(l: int, i: int, side: str):
    ln, inew = C.c_int64(), C.c_int64()
    if side not in ("left", "right"):
        raise ArgumentError("side must be :left or :right")
    if lib().asgo_child1(int(l), int(i), 1 if side == "right" else 0, C.byref(ln), C.byref(inew)) != 0:
        raise ArgumentError(f"invalid level-index pair ({l}, {i})")
    return ln.value, inew.value


def parent1(l: int, i: int):
    ln, inew = C.c_int64(), C.c_int64()
    if lib().asgo_parent1(int(l), int(i), C.byref(ln), C.byref(inew)) != 0:
        raise ArgumentError(f"invalid level-index pair ({l}, {i})")
    return ln.value, inew.value


def allindices_of_level(l: int) -> np.ndarray:
    n = lib().asgo_allindices_of_level(int(l), None)
    if n < 0:
        raise ArgumentError(f"invalid level {l}")
    out = np.empty(n, dtype=np.int64)
    lib().asgo_allindices_of_level(int(l), _i64(out))
    return out


def fraction2li(num: int, den: int):
    l, i = C.c_int64(), C.c_int64()
    if lib().asgo_fraction2li(int(num), int(den), C.byref(l), C.byref(i)) != 0:
        raise ArgumentError("denominator must be a power of two")
    return l.value, i.value


def allnodes1d(l: int):
    """src/math.jl:621-639"""
    if l < 1:
        raise ArgumentError(f"level must be >= 1 but got {l}")
    if l == 1:
        return [1], [1]
    if l == 2:
        return [1, 2, 2], [1, 0, 2]
    den = power2(l - 1)
    Ls, Is = [], []
    for i in range(0, den + 1):
        a, b = fraction2li(i, den)
        Ls.append(a)
        Is.append(b)
    return Ls, Is


def isbig2add(alpha, fx, rtol, atol, use_rtol) -> bool:
    return bool(lib().asgo_isbig2add(float(alpha), float(fx), float(rtol), float(atol), int(bool(use_rtol))))


def num_threads() -> int:
    return int(lib().asgo_num_threads())


# ---------------------------------------------------------------------------------------------
# node store + training logic
# ---------------------------------------------------------------------------------------------
class OracleGrid:
    """Restates ``SparseGridInterpolant{D}`` (src/class.jl:281-319). Arrays are row-major N x D."""

    def __init__(self, xdim: int, lb=None, ub=None):
        assert xdim > 0, "Dimension must be positive."
        lb = np.zeros(xdim) if lb is None else np.asarray(lb, dtype=np.float64)
        ub = np.ones(xdim) if ub is None else np.asarray(ub, dtype=np.float64)
        assert len(lb) == xdim and len(ub) == xdim
        assert np.all(ub > lb) and np.all(np.isfinite(lb)) and np.all(np.isfinite(ub))
        self.D = xdim
        self.lb = np.ascontiguousarray(lb, dtype=np.float64)
        self.ub = np.ascontiguousarray(ub, dtype=np.float64)
        self.levels = np.empty((0, xdim), dtype=np.int64)
        self.indices = np.empty((0, xdim), dtype=np.int64)
        self.depths = np.empty(0, dtype=np.int64)
        self.fvals = np.empty(0, dtype=np.float64)
        self.coefs = np.empty(0, dtype=np.float64)

    def __len__(self):
        return len(self.fvals)

    # -- src/class.jl:375-387 ------------------------------------------------------------------
    def stack(self, levels=None, indices=None) -> np.ndarray:
        Lv = np.ascontiguousarray(self.levels if levels is None else levels, dtype=np.int64)
        Ix = np.ascontiguousarray(self.indices if indices is None else indices, dtype=np.int64)
        N = Lv.shape[0]
        out = np.empty((N, self.D), dtype=np.float64)
        if N and lib().asgo_stack(self.D, N, _i64(Lv), _i64(Ix), self.D, 1, _f64(self.lb), _f64(self.ub), _f64(out)) != 0:
            raise ArgumentError("invalid level-index pair")
        return out

    # -- src/class.jl:629-680, 747-761 ---------------------------------------------------------
    def evaluate(self, X, max_depth: int | None = None, return_abssum: bool = False):
        X = np.ascontiguousarray(np.atleast_2d(np.asarray(X, dtype=np.float64)))
        assert X.shape[1] == self.D
        M = X.shape[0]
        out = np.zeros(M)
        ab = np.zeros(M)
        Lv = np.ascontiguousarray(self.levels)
        Ix = np.ascontiguousarray(self.indices)
        cf = np.ascontiguousarray(self.coefs)
        dp = np.ascontiguousarray(self.depths)
        rc = lib().asgo_interpolate_batch(
            self.D, len(cf), _i64(Lv), _i64(Ix), self.D, 1, _f64(cf), _f64(self.lb), _f64(self.ub),
            _f64(X), M, self.D, 1, _i64(dp) if max_depth is not None else None,
            int(max_depth) if max_depth is not None else 0, _f64(out), _f64(ab))
        if rc != 0:
            raise ArgumentError("invalid level-index pair")
        return (out, ab) if return_abssum else out

    def __call__(self, x):
        x = np.asarray(x, dtype=np.float64)
        if x.ndim == 1:
            return float(self.evaluate(x[None, :])[0])
        return self.evaluate(x)

    # -- src/class.jl:695-737 (documented intent: the reference's call site never reaches it, defect B3) --------
    def extrapolate(self, x) -> float:
        """First-order Taylor expansion at the nearest boundary point, one point, scalar loop as in the reference:
        f0 = G(xbnd); for every dimension j outside [0,1]: neighbour one ghost mesh step (src/math.jl:888-890,
        finest level of that dimension) towards the interior, one-sided slope (f0 - f(nbr)) / h, fval += slope *
        overshoot. Two deliberate readings of the documented intent ("1st order Taylor expansion"): lengths are
        unit-cube lengths (the reference rescales unit-cube points a second time, defect B4), and every slope is
        taken against f0 (the reference's loop uses the running fval, src/class.jl:731, which is a Taylor expansion
        only when a single coordinate is outside). With one coordinate outside on [0,1]^D this is the reference's
        arithmetic verbatim."""
        x = np.asarray(x, dtype=np.float64)
        w = self.ub - self.lb
        x01 = 0.0 + (1.0 * (x - self.lb)) / w                  # scale(), src/math.jl:308
        xb01 = np.minimum(np.maximum(x01, 0.0), 1.0)           # clamp.(x01, 0.0, 1.0)   (NaN stays NaN)
        to_x = lambda u: np.where(u >= 1.0, self.ub, np.where(u <= 0.0, self.lb, self.lb + w * u))
        f0 = fval = float(self.evaluate(to_x(xb01)[None, :])[0])
        maxl = self.levels.max(axis=0)
        for j in range(self.D):
            if 0.0 <= x01[j] <= 1.0:
                continue
            dx_out = x01[j] - xb01[j]
            h = 1.0 / power2(int(maxl[j]) - 1)                 # ghost_distance
            right = x01[j] > 1.0
            xn = xb01.copy()
            xn[j] -= h if right else -h
            fn = float(self.evaluate(to_x(xn)[None, :])[0])
            slope = ((f0 - fn) if right else (fn - f0)) / h
            fval += slope * dx_out
        return fval

    # -- src/class.jl:434-538 (documented intent) ----------------------------------------------
    def basis(self, X=None, dropzeros: bool = True, atol: float = 1e-16) -> np.ndarray:
        """Dense M x N matrix; entries the reference would drop are exactly 0.0."""
        X = self.stack() if X is None else np.ascontiguousarray(np.atleast_2d(np.asarray(X, dtype=np.float64)))
        M, N = X.shape[0], len(self)
        out = np.zeros((M, N))
        Lv = np.ascontiguousarray(self.levels)
        Ix = np.ascontiguousarray(self.indices)
        rc = lib().asgo_basis_dense(self.D, N, _i64(Lv), _i64(Ix), self.D, 1, _f64(self.lb), _f64(self.ub),
                                    _f64(X), M, self.D, 1, int(dropzeros), float(atol), _f64(out))
        if rc != 0:
            raise ArgumentError("invalid level-index pair")
        return out

    # -- src/train.jl:40-78 --------------------------------------------------------------------
    def rsg(self, maxdepth: int, maxlevels):
        ml = np.full(self.D, maxlevels, dtype=np.int64) if np.isscalar(maxlevels) else np.asarray(maxlevels, dtype=np.int64)
        if np.any(ml < 1):
            raise ArgumentError("maxlevels must be >= 1")
        if maxdepth < 2:
            raise ArgumentError("maxdepth must be >= 2")
        n = lib().asgo_rsg(self.D, maxdepth, _i64(ml), None, None, None)
        self.levels = np.empty((n, self.D), dtype=np.int64)
        self.indices = np.empty((n, self.D), dtype=np.int64)
        self.depths = np.empty(n, dtype=np.int64)
        lib().asgo_rsg(self.D, maxdepth, _i64(ml), _i64(self.levels), _i64(self.indices), _i64(self.depths))
        self.fvals = np.zeros(n)
        self.coefs = np.zeros(n)

    # -- src/train.jl:112-195 ------------------------------------------------------------------
    def train(self, f, maxdepth: int, maxlevels):
        self.rsg(maxdepth, maxlevels)
        for lnow in range(1, maxdepth + 1):
            idx = np.flatnonzero(self.depths == lnow)
            if lnow == 1:
                assert len(idx) == 1 and self.depths[0] == 1
                x = self.stack(np.ones((1, self.D), np.int64), np.ones((1, self.D), np.int64))[0]
                fv = float(f(x))
                self.fvals[0] = fv
                self.coefs[0] = fv
            elif len(idx):
                X = self.stack(self.levels[idx], self.indices[idx])
                fv = np.array([float(f(x)) for x in X])
                # coefficients of depth lnow are still zero and are masked out anyway (A11)
                g = self.evaluate(X, max_depth=lnow - 1)
                self.fvals[idx] = fv
                self.coefs[idx] = fv - g

    # -- src/train.jl:235-463 ------------------------------------------------------------------
    def adapt(self, f, maxdepth: int, tol: float = 1e-3, toltype: str = "absolute", log: list | None = None):
        assert len(self) > 0
        lnow = int(self.depths.max())
        assert 0 < lnow <= maxdepth
        if toltype == "absolute":
            atol, rtol = tol, float("nan")
        elif toltype == "relative":
            atol, rtol = float("nan"), tol
        else:
            raise ArgumentError("toltype must be :absolute or :relative.")
        D = self.D
        while lnow < maxdepth:
            parents = np.flatnonzero(self.depths == lnow)
            # candidates in the reference's visiting order: for j, for parent, left then right
            cl, ci = [], []
            for j in range(D):
                for p in parents:
                    for side in ("left", "right"):
                        L = self.levels[p].copy()
                        I = self.indices[p].copy()
                        L[j], I[j] = child1(int(L[j]), int(I[j]), side)
                        if isvalid(L, I):
                            cl.append(L)
                            ci.append(I)
            new = {}
            margin = float("inf")
            if cl:
                CL = np.array(cl, dtype=np.int64)
                CI = np.array(ci, dtype=np.int64)
                X = self.stack(CL, CI)
                fv = np.array([float(f(x)) for x in X])
                al = fv - self.evaluate(X)  # whole current grid (A12)
                for k in range(len(fv)):
                    if toltype == "absolute":
                        margin = min(margin, abs(abs(al[k]) - tol))
                    if isbig2add(al[k], fv[k], rtol, atol, toltype == "relative"):
                        key = (tuple(CL[k]), tuple(CI[k]))
                        if key not in new:
                            new[key] = (fv[k], al[k])
            if log is not None:
                log.append({"depth": lnow, "parents": len(parents), "candidates": len(cl), "added": len(new),
                            "min_margin": margin})
            if not new:
                break
            keys = list(new.keys())
            self.levels = np.vstack([self.levels, np.array([k[0] for k in keys], dtype=np.int64)])
            self.indices = np.vstack([self.indices, np.array([k[1] for k in keys], dtype=np.int64)])
            self.depths = np.concatenate([self.depths, np.array([sum(k[0]) - D + 1 for k in keys], dtype=np.int64)])
            self.fvals = np.concatenate([self.fvals, np.array([new[k][0] for k in keys])])
            self.coefs = np.concatenate([self.coefs, np.array([new[k][1] for k in keys])])
            lnow += 1
            if lnow == maxdepth:
                break


def refpath_eval(G: OracleGrid, X, nthreads: int = 0) -> np.ndarray:
    """Reference-structured CPU path (cost model of src/class.jl:635-642) for baseline timing."""
    X = np.ascontiguousarray(np.atleast_2d(np.asarray(X, dtype=np.float64)))
    Lc = np.asfortranarray(G.levels)
    Ic = np.asfortranarray(G.indices)
    cf = np.ascontiguousarray(G.coefs)
    out = np.zeros(X.shape[0])
    rc = lib().asgo_refpath_batch(G.D, len(cf), Lc.ctypes.data_as(c_i64p), Ic.ctypes.data_as(c_i64p), _f64(cf),
                                  _f64(G.lb), _f64(G.ub), _f64(X), X.shape[0], G.D, 1, _f64(out), int(nthreads))
    if rc != 0:
        raise ArgumentError("invalid level-index pair")
    return out
