"""Chained rollouts: n steps as n launches vs one fused launch (state on chip)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mbrl_traffic_b200.engine import Engine  # noqa: E402


def run(model, dtype, b, n, steps, fuse, fast=False, rewards=False, actions=None):
    os.environ["MT_NO_FUSE"] = "0" if fuse else "1"
    tdt = torch.float64 if dtype == "float64" else torch.float32
    eng = Engine(model, dtype, b, n)
    eng.set_params(dx=50, dt=0.5, rho_max=1, v_max=30, lam=1, tau=9,
                   boundary_conditions="loop", fast_math=fast)
    rng = np.random.default_rng(0)
    rho = rng.uniform(0.05, 0.9, (b, n))
    x = torch.as_tensor(np.concatenate((rho, 30 * (1 - rho) + rng.normal(0, 1, (b, n))), 1),
                        device="cuda").to(tdt)
    y = torch.empty_like(x)
    rew = torch.empty(steps, b, dtype=tdt, device="cuda") if rewards else None
    if actions is None:
        actions = rewards
    act = (torch.full((steps, b), 25.0, dtype=tdt, device="cuda") if actions else None)
    eng.rollout(x.clone(), y, steps, actions=act, rewards=rew)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a = x.clone()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.rollout(a, y, steps, actions=act, rewards=rew)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    eng.close()
    return b * n * steps / best


def main():
    out = []
    for model, dtype, b, n, steps, fast in [
            ("arz", "float64", 4096, 1000, 100, False),
            ("arz+actions+rewards", "float64", 4096, 1000, 100, False),
            ("arz+actions", "float64", 4096, 1000, 100, False),
            ("arz+rewards", "float64", 4096, 1000, 100, False),
            ("lwr+actions+rewards", "float64", 4096, 1000, 100, False),
            ("arz+actions+rewards", "float32", 4096, 1000, 100, True),
            ("arz", "float32", 4096, 1000, 100, True),
            ("lwr", "float64", 4096, 1000, 100, False),
            ("arz", "float64", 256, 1000, 200, False),
            ("lwr", "float64", 1, 20, 1000, False),
            ("arz", "float64", 1, 30, 1000, False)]:
        rw = "+rewards" in model                   # a reward row per step
        ac = "+actions" in model                   # per-step speed limits
        kind = model.split("+")[0]
        a = run(kind, dtype, b, n, steps, False, fast, rw, ac)
        f = run(kind, dtype, b, n, steps, True, fast, rw, ac)
        out.append(dict(model=model, dtype=dtype, fast=fast, envs=b, cells=n, steps=steps,
                        stepwise_gcups=round(a / 1e9, 3), fused_gcups=round(f / 1e9, 3),
                        speedup=round(f / a, 2)))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
