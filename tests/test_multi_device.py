"""Several GPUs behind the C ABI (asg_init_devices): one process, one replica of the packed grid
per device, host-buffer calls split their points over the devices. With one visible device the same code path runs
with a single replica; with two or more (gpurun --gpus N) the sharded results must equal the single-device ones
bit for bit -- the per-point arithmetic does not depend on which device evaluates the point."""
import subprocess
import sys
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_CHILD = r'''
import sys, json, ctypes as C
import numpy as np
sys.path.insert(0, %(root)r)
import asg_b200 as asg
from oracle import oracle as o
lib = asg._capi.load()
n = C.c_int32()
asg._capi.check(lib.asg_device_count(C.byref(n)))
ndev = int(%(want)d) if %(want)d > 0 else n.value
devs = list(range(ndev))
asg.init_devices(devs)
assert asg.context_devices() == devs
def f(x): return float(np.sum(np.sin(2 * np.pi * x)) + x[0] * x[1])
D = 5
G = asg.SparseGridInterpolant(D, lb=np.zeros(D), ub=np.arange(1, D + 1.0))
asg.train(G, f, 6, 5)                      # surplus_nodes per depth, alpha scattered to every replica
OG = o.OracleGrid(D, lb=np.zeros(D), ub=np.arange(1, D + 1.0))
OG.train(f, 6, 5)
assert np.array_equal(G.levels, OG.levels) and np.max(np.abs(G.coefs - OG.coefs)) < 1e-13
rng = np.random.default_rng(1)
M = 700001
X = rng.random((M, D)) * np.arange(1, D + 1.0)
X[::1000, 2] = -1.0
y = G.evaluate(X)                          # split over the devices (M >= 2^16 per device)
sel = rng.choice(M, size=3000, replace=False)
want, sc = OG.evaluate(X[sel], return_abssum=True)
err = np.abs(y[sel] - want)
assert np.all((err <= 1e-12 * np.abs(want)) | (err <= 1e-14 * np.maximum(1.0, sc)))
assert np.all(y[::1000] == 0.0)
fv = rng.standard_normal(M)
al = G.surplus(X, fv, max_depth=4)
wm = OG.evaluate(X[sel], max_depth=4)
assert np.max(np.abs(al[sel] - (fv[sel] - wm))) < 1e-12
# coefficient updates reach every replica
G.coefs[:] = 0.5 * G.coefs
asg._capi.check(lib.asg_grid_set_coefs(G._ensure(), 0, len(G), G.coefs.ctypes.data_as(asg._capi.f64p)))
y2 = G.evaluate(X)
assert np.array_equal(y2, 0.5 * y)
# adapt! on a fresh grid (append -> every replica re-uploaded)
G2 = asg.SparseGridInterpolant(D, lb=np.zeros(D), ub=np.arange(1, D + 1.0))
asg.train(G2, f, 6, 5); asg.adapt(G2, f, 8, tol=1e-3, verbose=False)
OG2 = o.OracleGrid(D, lb=np.zeros(D), ub=np.arange(1, D + 1.0)); OG2.train(f, 6, 5); OG2.adapt(f, 8, tol=1e-3)
w2, s2 = OG2.evaluate(X[sel], return_abssum=True)
e2 = np.abs(G2.evaluate(X)[sel] - w2)
assert np.all((e2 <= 1e-12 * np.abs(w2)) | (e2 <= 1e-14 * np.maximum(1.0, s2)))
np.save(%(out)r, y)
print(json.dumps({"ndev": ndev, "launches": asg.launch_count()}))
'''


def _run(want, out):
    code = _CHILD % {"root": ROOT, "want": want, "out": out}
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout.strip().splitlines()[-1]


def test_sharded_equals_single_device(tmp_path):
    """all visible devices vs device 0 alone, in separate processes (one context per process)"""
    import json
    a = str(tmp_path / "all.npy")
    b = str(tmp_path / "one.npy")
    info_all = json.loads(_run(0, a))
    info_one = json.loads(_run(1, b))
    assert info_one["ndev"] == 1 and info_all["ndev"] >= 1
    ya, yb = np.load(a), np.load(b)
    assert np.array_equal(ya, yb)            # bit-identical whatever the number of devices
