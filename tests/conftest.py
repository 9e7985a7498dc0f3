import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RTOL = 1e-12   # BASELINE.json north_star: fp64 values within 1e-12 relative ...
ATOL = 1e-14   # ... 1e-14 absolute near zero (scaled by the magnitude of the summed terms)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def assert_close(got, want, scale=None, what=""):
    """|got - want| <= 1e-12 |want|  or  <= 1e-14 * max(1, sum_n |c_n phi_n|)."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    s = np.ones_like(want) if scale is None else np.maximum(1.0, np.asarray(scale, dtype=np.float64))
    err = np.abs(got - want)
    ok = (err <= RTOL * np.abs(want)) | (err <= ATOL * s)
    both_nan = np.isnan(got) & np.isnan(want)
    ok |= both_nan
    if not np.all(ok):
        k = int(np.argmax(np.where(ok, 0.0, err)))
        raise AssertionError(f"{what}: {np.count_nonzero(~ok)} of {ok.size} values out of tolerance; worst at {k}: "
                             f"got {got.flat[k]!r} want {want.flat[k]!r} err {err.flat[k]:.3e}")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def asg():
    import asg_b200
    return asg_b200


def sinsum(x):
    return float(np.sum(np.sin(2 * np.pi * np.asarray(x))))


def utility(x):
    return -1.0 / (0.1 + float(np.sum(x)))


def parabola(x):
    x = np.asarray(x)
    return float(np.prod(4 * x * (1 - x)))
