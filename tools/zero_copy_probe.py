"""Probe: the step kernel reading its input tiles straight from page-locked
HOST memory and writing its results straight back (bulk copies over PCIe,
no staging copies), against the staged host path (mt_step_host)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mbrl_traffic_b200.engine import Engine, PinnedArray  # noqa: E402

b, n = 4096, 1000
eng = Engine("arz", "float64", b, n, host_staging=True)
eng.set_params(dx=50, dt=0.5, rho_max=1, v_max=30, lam=1, tau=9,
               boundary_conditions="loop", stream=0)
rng = np.random.default_rng(0)
rho = rng.uniform(0.05, 0.9, (b, n))
pin_in = PinnedArray((b, 2 * n), np.float64)
pin_out = PinnedArray((b, 2 * n), np.float64)
pin_ref = PinnedArray((b, 2 * n), np.float64)
pin_in.array[...] = np.concatenate((rho, 30 * (1 - rho) + rng.normal(0, 1, (b, n))), 1)
dev_in = torch.as_tensor(pin_in.array, device="cuda")
dev_out = torch.empty_like(dev_in)
s = torch.cuda.Stream()
res = {}


def timeit(fn, k=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        for _ in range(k):
            fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) / k)
    return float(np.median(ts))


t = timeit(lambda: eng.step_host(pin_in.array, pin_ref.array))
res["staged_ms"] = t * 1e3
hin, hout = pin_in.array.ctypes.data, pin_out.array.ctypes.data
for name, a, o in (("host_to_host", hin, hout), ("host_to_dev", hin, dev_out.data_ptr()),
                   ("dev_to_host", dev_in.data_ptr(), hout)):
    def fn():
        eng.step(a, o, stream=s)
        s.synchronize()
    try:
        t = timeit(fn)
        res[name + "_ms"] = t * 1e3
        res[name + "_GBps"] = b * 2 * n * 8 / t / 1e9
    except Exception as exc:
        res[name + "_error"] = repr(exc)[:200]
res["bit_equal_to_staged"] = bool(np.array_equal(pin_out.array, pin_ref.array))
print(json.dumps(res))
