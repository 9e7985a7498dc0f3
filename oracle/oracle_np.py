"""Independent numpy restatement of the evaluation path -- TEST INFRASTRUCTURE ONLY.

Second, deliberately separate restatement of src/math.jl:332-357 and src/class.jl:629-664, written
against the reference text (not against asg_oracle.c) so the two can be checked against each other.
Terms are formed in double exactly as the reference does, ``((c*phi_1)*phi_2)...`` and then summed
with ``math.fsum`` (exact), which is the ground truth the C oracle's long-double sum approximates.
Use only for small cases: it is O(M*N*D) Python/numpy work.
"""
from __future__ import annotations

import math

import numpy as np


def phi_vec(x: float, L: np.ndarray, I: np.ndarray) -> np.ndarray:
    """phi.(x, L, I) for one coordinate x against N (level, index) pairs -- src/math.jl:345-357."""
    L = np.asarray(L, dtype=np.int64)
    I = np.asarray(I, dtype=np.int64)
    out = np.empty(L.shape, dtype=np.float64)
    hi = L > 2
    if np.any(hi):
        # power2(l-1) = 2 << (l-2)  (src/math.jl:59-61)
        s = np.ldexp(1.0, (L[hi] - 1).astype(np.int64))
        t = x * s - I[hi].astype(np.float64)
        with np.errstate(invalid="ignore"):
            inside = (-1.0 <= t) & (t <= 1.0)
        out[hi] = np.where(inside, 1.0 - np.abs(t), 0.0)
    m = (L == 2) & (I == 0)
    out[m] = (1.0 - 2.0 * x) if (0 <= x <= 0.5) else 0.0
    m2 = (L == 2) & (I == 2)
    out[m2] = (2.0 * x - 1.0) if (0.5 <= x <= 1) else 0.0
    m1 = L == 1
    out[m1] = 1.0 if (0 <= x <= 1) else 0.0
    bad = ~(hi | m | m2 | m1)
    if np.any(bad):
        raise ValueError("invalid level-index pair")
    return out


def scale01(x: np.ndarray, lb: np.ndarray, ub: np.ndarray) -> np.ndarray:
    """src/math.jl:301-309 with to_lb = 0, to_ub = 1."""
    x = np.asarray(x, dtype=np.float64)
    return 0.0 + (1.0 - 0.0) * (x - lb) / (ub - lb)


def terms(levels, indices, coefs, lb, ub, x, mask=None) -> np.ndarray:
    x01 = scale01(x, np.asarray(lb, float), np.asarray(ub, float))
    if mask is None:
        mask = np.ones(len(coefs), dtype=bool)
    vals = np.array(coefs, dtype=np.float64)[mask]
    Lm = np.asarray(levels)[mask]
    Im = np.asarray(indices)[mask]
    for d in range(Lm.shape[1]):
        vals = vals * phi_vec(float(x01[d]), Lm[:, d], Im[:, d])
    return vals


def interpolate(levels, indices, coefs, lb, ub, x, mask=None) -> float:
    """src/class.jl:629-664; exact sum of the double-rounded terms."""
    return math.fsum(terms(levels, indices, coefs, lb, ub, x, mask))


def basis_row(levels, indices, lb, ub, x, dropzeros=True, atol=1e-16) -> np.ndarray:
    v = terms(levels, indices, np.ones(len(levels)), lb, ub, x)
    if dropzeros:
        v = np.where(v < atol, 0.0, v)
    return v
