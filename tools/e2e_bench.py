"""End-to-end host path: get_next_obs on pinned host arrays (mt_step_host),
plus raw pinned copy bandwidth for reference."""
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
pin_in.array[...] = np.concatenate((rho, 30 * (1 - rho)), 1)
for _ in range(3):
    eng.step_host(pin_in.array, pin_out.array)
t0 = time.perf_counter()
k = 20
for _ in range(k):
    eng.step_host(pin_in.array, pin_out.array)
dt = (time.perf_counter() - t0) / k
x = torch.empty(b, 2 * n, dtype=torch.float64, device="cuda")
h = torch.from_numpy(pin_in.array)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    x.copy_(h, non_blocking=True)
torch.cuda.synchronize()
h2d = (time.perf_counter() - t0) / 10
ho = torch.from_numpy(pin_out.array)
t0 = time.perf_counter()
for _ in range(10):
    ho.copy_(x, non_blocking=True)
torch.cuda.synchronize()
d2h = (time.perf_counter() - t0) / 10
# both directions at once (two streams): the ceiling of the pipelined host path
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
y = torch.empty_like(x)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1):
        x.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        ho.copy_(y, non_blocking=True)
torch.cuda.synchronize()
bidir = (time.perf_counter() - t0) / 10
print(json.dumps(dict(chunk_mb=os.environ.get("MT_HOST_CHUNK_MB", "8"),
                      ms_per_step=round(dt * 1e3, 3),
                      gcups=round(b * n / dt / 1e9, 3),
                      h2d_GBps=round(x.numel() * 8 / h2d / 1e9, 1),
                      d2h_GBps=round(x.numel() * 8 / d2h / 1e9, 1),
                      bidir_GBps_each=round(x.numel() * 8 / bidir / 1e9, 1),
                      bidir_ms=round(bidir * 1e3, 3))))
