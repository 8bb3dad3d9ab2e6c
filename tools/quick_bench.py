"""Developer timing loop for the step kernels (not the contract bench).

    python tools/quick_bench.py [--model arz] [--dtype float64] [--envs 4096]
                                [--cells 1000] [--sets 4] [--steps 200]

Round-robins over ``--sets`` independent batches so that every step reads its
input from HBM (working set >> L2), and also reports the chained
(ping-pong on one batch) rate.
"""
import argparse
import json
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mbrl_traffic_b200.engine import Engine  # noqa: E402

BYTES = {"arz": 4, "lwr": 3, "nonlocal": 3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="arz")
    ap.add_argument("--dtype", default="float64")
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--cells", type=int, default=1000)
    ap.add_argument("--sets", type=int, default=4)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--fast", action="store_true")
    ap.add_argument("--reward", action="store_true")
    ap.add_argument("--direct", action="store_true")
    ap.add_argument("--n-eta", type=int, default=16)
    ap.add_argument("--per-env", action="store_true",
                    help="pass the parameters as per-road device arrays")
    ap.add_argument("--rho-max", type=float, default=1.0)
    ap.add_argument("--independent", action="store_true",
                    help="MT_FLAG_INDEPENDENT_INPUT (valid for the streaming measurement only)")
    ap.add_argument("--lam", type=float, default=1.0)
    ap.add_argument("--scheme", default="godunov", help="non-local model: godunov | lxf")
    ap.add_argument("--action", action="store_true", help="pass a per-road speed limit")
    ap.add_argument("--bc", default="loop",
                    help="loop | extend_both | inflow_outflow (constant 0.3 rho_max inflow, "
                         "free outflow: BASELINE config 3)")
    ap.add_argument("--population", action="store_true",
                    help="one random parameter vector per road (an optimizer's population)")
    args = ap.parse_args()

    tdt = torch.float64 if args.dtype == "float64" else torch.float32
    b, n = args.envs, args.cells
    eng = Engine(args.model, args.dtype, b, n, max_n_eta=256)
    gamma = None
    if args.model == "nonlocal":
        from mbrl_traffic_b200.nonlocal_ import lookahead_weights
        gamma = torch.as_tensor(lookahead_weights(50.0 * args.n_eta, 50.0),
                                device="cuda").to(tdt)
    def par(x):
        if not args.per_env:
            return x
        return torch.full((b,), float(x), dtype=tdt, device="cuda")
    bc = {"inflow_outflow": 0.3 * args.rho_max} if args.bc == "inflow_outflow" else args.bc
    if args.population:
        g = torch.Generator(device="cuda").manual_seed(1)
        def rnd(lo, hi):
            return lo + (hi - lo) * torch.rand(b, generator=g, dtype=tdt, device="cuda")
        eng.set_params(dx=50, dt=0.5, rho_max=rnd(0.9, 1.5) * args.rho_max, v_max=rnd(20, 40),
                       lam=rnd(0.5, 3.0), boundary_conditions=bc, tau=rnd(5, 25),
                       fast_math=args.fast, no_pipeline=args.direct, gamma=gamma,
                       independent_input=args.independent, scheme=args.scheme)
    else:
        eng.set_params(dx=50, dt=0.5, rho_max=par(args.rho_max), v_max=par(30),
                       lam=par(args.lam), boundary_conditions=bc, tau=par(9),
                       fast_math=args.fast, no_pipeline=args.direct, gamma=gamma,
                       independent_input=args.independent, scheme=args.scheme)
    rng = np.random.default_rng(0)
    rho = args.rho_max * rng.uniform(0.05, 0.9, (b, n))
    v = 30 * (1 - rho / args.rho_max) + rng.normal(0, 1, (b, n))
    obs = torch.as_tensor(np.concatenate((rho, v), 1), device="cuda").to(tdt)
    ins = [obs.clone() for _ in range(args.sets)]
    outs = [torch.empty_like(obs) for _ in range(args.sets)]
    rew = torch.empty(b, dtype=tdt, device="cuda") if args.reward else None
    act = torch.full((b,), 25.0, dtype=tdt, device="cuda") if args.action else None
    esz = obs.element_size()
    cells = b * n

    def run(k, chained):
        if chained:
            a, c = ins[0], outs[0]
            for _ in range(k):
                eng.step(a, c, reward=rew, action=act)
                a, c = c, a
        else:
            for i in range(k):
                j = i % args.sets
                eng.step(ins[j], outs[j], reward=rew, action=act)

    res = {}
    side = torch.cuda.Stream()
    for chained in (False, True):
        # the K launches are captured once in a CUDA graph and replayed, so the
        # numbers are free of Python / ctypes launch overhead
        with torch.cuda.stream(side):
            run(args.warmup, chained)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                run(args.steps, chained)
        torch.cuda.synchronize()
        graph.replay()
        torch.cuda.synchronize()
        best = None
        for _ in range(3):
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            best = ms if best is None else min(best, ms)
        cu = cells / (best * 1e-3)
        gbs = cu * BYTES[args.model] * esz / 1e9
        res["chained" if chained else "streaming"] = dict(
            ms_per_step=round(best, 5), gcups=round(cu / 1e9, 2),
            algo_GBps=round(gbs, 1), frac_of_6545=round(gbs / 6545.6, 3))
        if chained:
            ins[0].copy_(obs)
    print(json.dumps(dict(model=args.model, dtype=args.dtype, envs=b, cells=n,
                          fast=args.fast, reward=args.reward, direct=args.direct,
                          rho_max=args.rho_max, **res)))


if __name__ == "__main__":
    main()
