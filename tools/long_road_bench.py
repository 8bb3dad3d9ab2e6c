"""Single long road (config 4 shape on one GPU): cell-updates/s of the step
for a 10 M-cell ARZ highway, pipelined segments vs the direct kernel."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mbrl_traffic_b200.engine import Engine  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    rng = np.random.default_rng(0)
    rho = rng.uniform(0.05, 0.9, n)
    v = 30 * (1 - rho) + rng.normal(0, 1, n)
    x = torch.as_tensor(np.concatenate((rho, v)), device="cuda").reshape(1, -1)
    y = torch.empty_like(x)
    out = {"cells": n}
    for name, direct in (("segments", False), ("direct", True)):
        eng = Engine("arz", "float64", 1, n)
        eng.set_params(dx=50, dt=0.5, rho_max=1, v_max=30, lam=1, tau=9,
                       boundary_conditions="extend_both", no_pipeline=direct)
        for _ in range(5):
            eng.step(x, y)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        k = 50
        for i in range(k):
            eng.step(x, y) if i % 2 == 0 else eng.step(y, x)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / k
        out[name] = dict(us_per_step=round(us, 2), gcups=round(n / us / 1e3, 2),
                         frac_of_6545=round(n * 32 / us / 1e3 / 6545.6, 3))
        eng.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
