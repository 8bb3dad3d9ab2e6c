"""Small launches of every kernel path added in round 2, for compute-sanitizer
(one tool per run: `compute-sanitizer --tool racecheck|memcheck python
tools/sanitize_smoke.py`).  Sizes are small (the tools slow kernels down 10-100x)
but keep the shapes that select each path: 1000-cell roads (4 cells per thread,
roads wider than a warp), more tiles than CTAs are not needed."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mbrl_traffic_b200.engine import Engine, actor_head  # noqa: E402
from mbrl_traffic_b200.nonlocal_ import lookahead_weights  # noqa: E402

P = dict(dx=50, dt=0.5, rho_max=1, v_max=30, lam=1, tau=9)


def obs(b, n, seed=0):
    rng = np.random.default_rng(seed)
    rho = rng.uniform(0.05, 0.9, (b, n))
    v = 30 * (1 - rho) + rng.normal(0, 1, (b, n))
    return torch.as_tensor(np.concatenate((rho, v), 1), device="cuda")


def main():
    b, n = 40, 1000
    x = obs(b, n)
    y = torch.empty_like(x)
    rew = torch.empty(b, dtype=torch.float64, device="cuda")
    act = torch.full((b,), 25.0, dtype=torch.float64, device="cuda")
    o16 = torch.empty(b, 2 * n, dtype=torch.bfloat16, device="cuda")
    e = Engine("arz", "float64", b, n)
    e.set_params(boundary_conditions="loop", **P)
    e.step(x, y)                                   # plain step
    e.step(x, y, reward=rew)                       # reducer warp
    e.step(x, y, reward=rew, action=act, obs16=o16)
    acts = torch.full((6, b), 25.0, dtype=torch.float64, device="cuda")
    rews = torch.empty(6, b, dtype=torch.float64, device="cuda")
    e.rollout(x.clone(), y, 6, actions=acts, rewards=rews)      # fused, reducer warp per step
    e.rollout(x.clone(), y, 6)
    e.set_params(boundary_conditions="loop", fsolve_abort=True, **P)
    z = x.clone()
    z[3, 17] = 0.0
    e.step(z, y, reward=rew, obs16=o16)            # abort fix-up
    e.close()
    t = torch.float64
    e = Engine("arz", "float64", b, n)             # per-road table: staged rows
    e.set_params(dx=50, dt=0.5, rho_max=torch.full((b,), 1.0, dtype=t, device="cuda"),
                 v_max=torch.full((b,), 30.0, dtype=t, device="cuda"),
                 lam=torch.full((b,), 1.0, dtype=t, device="cuda"),
                 tau=torch.full((b,), 9.0, dtype=t, device="cuda"), boundary_conditions="loop")
    e.step(x, y, reward=rew)
    e.close()
    e = Engine("lwr", "float64", b, n)
    e.set_params(dx=50, dt=0.5, rho_max=1, v_max=30, lam=1, boundary_conditions="loop")
    e.step(x, y, reward=rew, obs16=o16)
    e.rollout(x.clone(), y, 4, actions=acts[:4].contiguous(), rewards=rews[:4].contiguous())
    e.close()
    g = torch.as_tensor(lookahead_weights(800.0, 50.0, "linear"), device="cuda")
    for scheme in ("godunov", "lxf"):
        e = Engine("nonlocal", "float64", b, n, max_n_eta=64)
        e.set_params(dx=50, dt=0.5, rho_max=1, v_max=30, lam=1,
                     boundary_conditions={"inflow_outflow": 0.3}, gamma=g, scheme=scheme)
        e.step(x, y, reward=rew)
        e.close()
    e = Engine("arz", "float64", 1, 40000)         # long road: windows, several steps per launch
    e.set_params(boundary_conditions="extend_both", **P)
    xr = obs(1, 40000, 1)
    e.rollout(xr.clone(), torch.empty_like(xr), 8)
    e.close()
    h = torch.randn(b, 256, device="cuda").relu().to(torch.bfloat16)
    w = torch.randn(1, 256, device="cuda").to(torch.bfloat16)
    bias = torch.zeros(1, device="cuda").to(torch.bfloat16)
    actor_head(h, w, bias, act, 15.0)
    torch.cuda.synchronize()
    print("sanitize_smoke ok")


if __name__ == "__main__":
    main()
