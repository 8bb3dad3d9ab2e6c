"""Launches every kernel family DESIGN.md section 5 quotes, a few times each,
so that ONE `ncu --set full` call captures them all (developer tool; the
summaries under profiles/ are cut from its report).

    python tools/profile_kernels.py [--launches 3] [--only name,name]

    ncu --set full --clock-control none --import-source on \
        -k regex:'step_|nonlocal' -o gpurun_out/r02_kernels python tools/profile_kernels.py

Cases (each prints "case <name>: <kernel launches>" so the report's launch
order can be mapped back):
  arz_f64            step_pipe_kernel<double, ARZ>, scalar parameters, 4096 x 1000
  arz_f64_table      the same batch, per-road parameter table
  arz_f64_reward     scalar parameters + per-road reward
  arz_f64_action     scalar parameters + per-road speed limit + reward (config 5's step)
  lwr_f64            step_pipe_kernel<double, LWR>, 16384 x 1000 (dynamic tile schedule)
  arz_f32_exact      step_ring_kernel<float, ARZ>, 16384 x 1000
  arz_f32_fast       step_pipe_kernel<float, ARZ, FAST>, 16384 x 1000
  nonlocal_f64       nonlocal_pipe_kernel<double>, 16384 x 1000, n_eta = 16, inflow / outflow
  seg_f64            step_seg_kernel<double, ARZ>: one step of a 10 M-cell road
  segfused_f64       step_segfused_kernel<double, ARZ>: 16 steps of that road in one launch
  fused_f64          step_fused_kernel<double, ARZ>: 20 steps of 4096 x 1000 in one launch
  fused_f64_reward   the same with a reward row per step and per-step speed limits
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mbrl_traffic_b200.engine import Engine  # noqa: E402

P = dict(dx=50, dt=0.5, rho_max=1, v_max=30, lam=1, tau=9)


def obs(b, n, dtype, seed=0):
    rng = np.random.default_rng(seed)
    rho = rng.uniform(0.05, 0.9, (b, n))
    v = 30 * (1 - rho) + rng.normal(0, 1, (b, n))
    t = torch.float64 if dtype == "float64" else torch.float32
    return torch.as_tensor(np.concatenate((rho, v), 1), device="cuda").to(t)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--launches", type=int, default=3)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    only = set(x for x in args.only.split(",") if x)
    k = args.launches

    def want(name):
        return not only or name in only

    def steps(name, eng, x, **kw):
        """k launches over 2 buffers (input never the previous output)."""
        ins = [x, x.clone()]
        outs = [torch.empty_like(x), torch.empty_like(x)]
        n0 = eng.launch_count
        for i in range(k):
            eng.step(ins[i % 2], outs[i % 2], **kw)
        torch.cuda.synchronize()
        print("case %s: %d launches" % (name, eng.launch_count - n0), flush=True)

    def tensor(b, val, t):
        return torch.full((b,), float(val), dtype=t, device="cuda")

    b, n = 4096, 1000
    if want("arz_f64"):
        e = Engine("arz", "float64", b, n)
        e.set_params(boundary_conditions="loop", **P)
        steps("arz_f64", e, obs(b, n, "float64"))
        e.close()
    if want("arz_f64_table"):
        e = Engine("arz", "float64", b, n)
        t = torch.float64
        e.set_params(dx=50, dt=0.5, rho_max=tensor(b, 1, t), v_max=tensor(b, 30, t),
                     lam=tensor(b, 1, t), tau=tensor(b, 9, t), boundary_conditions="loop")
        steps("arz_f64_table", e, obs(b, n, "float64"))
        e.close()
    if want("arz_f64_reward"):
        e = Engine("arz", "float64", b, n)
        e.set_params(boundary_conditions="loop", **P)
        steps("arz_f64_reward", e, obs(b, n, "float64"),
              reward=torch.empty(b, dtype=torch.float64, device="cuda"))
        e.close()
    if want("arz_f64_action"):
        e = Engine("arz", "float64", b, n)
        e.set_params(boundary_conditions="loop", **P)
        steps("arz_f64_action", e, obs(b, n, "float64"),
              action=tensor(b, 25, torch.float64),
              reward=torch.empty(b, dtype=torch.float64, device="cuda"))
        e.close()
    b = 16384
    if want("lwr_f64"):
        e = Engine("lwr", "float64", b, n)
        e.set_params(boundary_conditions="loop", **P)
        steps("lwr_f64", e, obs(b, n, "float64"))
        e.close()
    for name, fast in (("arz_f32_exact", False), ("arz_f32_fast", True)):
        if want(name):
            e = Engine("arz", "float32", b, n)
            e.set_params(boundary_conditions="loop", fast_math=fast, **P)
            steps(name, e, obs(b, n, "float32"))
            e.close()
    if want("nonlocal_f64"):
        from mbrl_traffic_b200.nonlocal_ import lookahead_weights
        g = torch.as_tensor(lookahead_weights(50.0 * 16, 50.0, "linear"), device="cuda")
        e = Engine("nonlocal", "float64", b, n, max_n_eta=256)
        e.set_params(dx=50, dt=0.5, rho_max=1, v_max=30, lam=1,
                     boundary_conditions={"inflow_outflow": 0.3}, gamma=g)
        steps("nonlocal_f64", e, obs(b, n, "float64"))
        e.close()
    n_road = 10_000_000
    if want("seg_f64") or want("segfused_f64"):
        e = Engine("arz", "float64", 1, n_road)
        e.set_params(boundary_conditions="extend_both", **P)
        x = obs(1, n_road, "float64")
        if want("seg_f64"):
            steps("seg_f64", e, x)
        if want("segfused_f64"):
            y = torch.empty_like(x)
            n0 = e.launch_count
            for _ in range(k):
                e.rollout(x.clone(), y, 16)
            torch.cuda.synchronize()
            print("case segfused_f64: %d launches" % (e.launch_count - n0), flush=True)
        e.close()
    b = 4096
    if want("fused_f64") or want("fused_f64_reward"):
        e = Engine("arz", "float64", b, n)
        e.set_params(boundary_conditions="loop", **P)
        x = obs(b, n, "float64")
        y = torch.empty_like(x)
        if want("fused_f64"):
            n0 = e.launch_count
            for _ in range(k):
                e.rollout(x.clone(), y, 20)
            torch.cuda.synchronize()
            print("case fused_f64: %d launches" % (e.launch_count - n0), flush=True)
        if want("fused_f64_reward"):
            acts = torch.full((20, b), 25.0, dtype=torch.float64, device="cuda")
            rew = torch.empty(20, b, dtype=torch.float64, device="cuda")
            n0 = e.launch_count
            for _ in range(k):
                e.rollout(x.clone(), y, 20, actions=acts, rewards=rew)
            torch.cuda.synchronize()
            print("case fused_f64_reward: %d launches" % (e.launch_count - n0), flush=True)
        e.close()


if __name__ == "__main__":
    main()
