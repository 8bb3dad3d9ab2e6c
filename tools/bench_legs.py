"""The other BASELINE.json configurations, measured beside the headline.

bench.py runs this file in a child process (one rank per GPU under torchrun
when N > 1) so that a fault here cannot take the headline measurement down.
One JSON line per leg on stdout (rank 0):

  nonlocal   config 3: non-local look-ahead model, 16384 roads x 1000 cells per
             GPU, n_eta = 16, constant inflow / free outflow, road-sharded
  rollout    config 5: MLP actor [2N -> 256 -> 256 -> 1] driving 4096 x 1000
             ARZ roads per GPU for a 500-step horizon (actor + step + reward)
  long_road  config 4: ONE 10 M-cell ARZ road; N = 1: one GPU; N > 1: 1-D
             domain decomposition, ghost cells exchanged over peer memory;
             checked cell by cell against the single-GPU run

Times are device times (CUDA events), max over ranks; rates are whole-job.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PEAK_FALLBACK = 6650.0


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return PEAK_FALLBACK


class Ctx(object):
    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.dist = None
        if self.world > 1:
            import datetime
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev,
                                    timeout=datetime.timedelta(seconds=120))
            self.dist = dist
        self.peak = measured_peak()

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def slowest(self, seconds):
        if self.dist is None:
            return seconds
        t = torch.tensor([seconds], dtype=torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_ok(self, ok):
        if self.dist is None:
            return bool(ok)
        t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item())

    def emit(self, obj):
        if self.rank == 0:
            print(json.dumps(obj), flush=True)


def timed_replays(ctx, graph, reps, stream):
    """Median device time of ``reps`` replays, each bracketed by a barrier."""
    times = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        ctx.barrier()
        with torch.cuda.stream(stream):
            graph.replay()                      # untimed: the GPU idled at the barrier
            e0.record(stream)
            graph.replay()
            e1.record(stream)
        ctx.barrier()
        times.append(ctx.slowest(e0.elapsed_time(e1) * 1e-3))
    return float(np.median(times)), times


def make_obs(b, n, seed, rho_hi=0.9):
    rng = np.random.default_rng(1234 + seed)
    rho = rng.uniform(0.05, rho_hi, (b, n))
    v = 30.0 * (1.0 - rho) + rng.normal(0.0, 1.0, (b, n))
    return np.concatenate((rho, v), axis=1)


# --------------------------------------------------------------------------- #
def leg_nonlocal(ctx, args):
    from mbrl_traffic_b200.engine import Engine
    from mbrl_traffic_b200.nonlocal_ import lookahead_weights
    b, n, n_eta = args.nonlocal_envs, 1000, 16
    dx, dt = 50.0, 0.5
    gamma_np = lookahead_weights(dx * n_eta, dx, "linear")
    gamma = torch.as_tensor(gamma_np, device=ctx.dev)
    bc = {"inflow_outflow": 0.3}
    eng = Engine("nonlocal", "float64", b, n, device=ctx.local_rank, max_n_eta=256)
    sets = 2
    ins = [torch.as_tensor(make_obs(b, n, 100 + ctx.rank * sets + j), device=ctx.dev)
           for j in range(sets)]
    outs = [torch.empty_like(x) for x in ins]
    stream = torch.cuda.Stream(device=ctx.dev)
    steps = args.nonlocal_steps

    def measure(scheme):
        eng.set_params(dx=dx, dt=dt, rho_max=1, v_max=30, lam=1, boundary_conditions=bc,
                       gamma=gamma, scheme=scheme)
        with torch.cuda.stream(stream):
            for k in range(4):
                eng.step(ins[k % sets], outs[k % sets], stream=stream)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                for k in range(steps):
                    eng.step(ins[k % sets], outs[k % sets], stream=stream)
        graph.replay()
        torch.cuda.synchronize()
        sec_, _ = timed_replays(ctx, graph, 5, stream)
        ok = ctx.all_ok(bool(torch.isfinite(outs[0]).all().item()))
        return sec_, ok

    # the second scheme (Lax-Friedrichs, mean downstream density) rides along
    sec_lxf, finite_lxf = measure("lxf")
    sec, finite = measure("godunov")
    # (parity of these kernels: tests/test_gpu_parity.py::test_nonlocal_*,
    #  tests/test_nonlocal_unpinned.py)
    cells = b * n
    per_step = sec / steps
    gbs = cells * 24 / per_step / 1e9
    ctx.emit({"leg": "nonlocal",
              "workload": "non-local look-ahead model, %d roads x %d cells per GPU, n_eta=%d "
                          "(linear kernel), constant inflow 0.3 / free outflow, float64 "
                          "(BASELINE.json configs[2])" % (b, n, n_eta),
              "value": ctx.world * cells / per_step, "unit": "cell-updates/s",
              "n_gpus": ctx.world, "ms_per_step": per_step * 1e3, "steps": steps, "reps": 5,
              "algorithmic_bytes_per_cell_update": 24,
              "roofline_frac": gbs / ctx.peak, "achieved_GBps": gbs,
              "scaling": "weak", "outputs_finite": finite,
              "lxf": {"value": ctx.world * cells / (sec_lxf / steps), "unit": "cell-updates/s",
                      "roofline_frac": cells * 24 / (sec_lxf / steps) / 1e9 / ctx.peak,
                      "outputs_finite": finite_lxf,
                      "what": "the same batch under scheme=\"lxf\" (Blandin-Goatin model, "
                              "Lax-Friedrichs flux)"},
              "parity_note": "parity unpinned: no reference implementation exists; the oracle "
                             "restates the cited paper (tests/test_gpu_parity.py::test_nonlocal_*)",
              "l2": "2 independent batches round-robin, %.0f MB working set per GPU"
                    % (2 * sets * cells * 16 / 1e6)})
    eng.close()


# --------------------------------------------------------------------------- #
def leg_rollout(ctx, args):
    from mbrl_traffic_b200 import ARZModel, ARZ_MODEL_PARAMS
    from mbrl_traffic_b200.rollout import BatchedMacroEnv, FeedForwardActor, PolicyRollout
    b, n, horizon = args.rollout_envs, 1000, args.horizon
    p = dict(ARZ_MODEL_PARAMS)
    p.pop("optimizer_cls")
    model = ARZModel(None, None, None, None, 0, optimizer_cls=None,
                     action_is_speed_limit=True, device=ctx.local_rank, **p)
    obs = torch.as_tensor(make_obs(b, n, 200 + ctx.rank), device=ctx.dev)
    env = BatchedMacroEnv(model, b, n, horizon=horizon, dtype=torch.float64,
                          obs_bf16=not args.rollout_plain)
    torch.manual_seed(1)          # (a seed whose random-init actor keeps the scheme stable)
    actor = FeedForwardActor(2 * n).to(ctx.dev).to(torch.bfloat16)
    runner = PolicyRollout(env, actor)     # captured once, replayed for every rollout
    runner.run(obs, horizon=8)
    torch.cuda.synchronize()
    times = []
    for _ in range(3):
        ctx.barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        total = runner.run(obs)
        e1.record()
        ctx.barrier()
        times.append(ctx.slowest(e0.elapsed_time(e1) * 1e-3))
    sec = float(np.median(times))
    finite = ctx.all_ok(bool(torch.isfinite(total).all().item()))
    ctx.emit({"leg": "rollout",
              "workload": "MLP actor [2N->256->256->1] (bf16) driving %d x %d ARZ roads per GPU, "
                          "horizon %d, speed-limit actions, density-weighted mean-speed reward, "
                          "float64 state (BASELINE.json configs[4])" % (b, n, horizon),
              "value": ctx.world * b * n * horizon / sec, "unit": "cell-updates/s",
              "road_steps_per_s": ctx.world * b * horizon / sec,
              "n_gpus": ctx.world, "seconds": sec, "us_per_step": sec / horizon * 1e6,
              "scaling": "weak", "returns_finite": finite, "reps": 3,
              "policy_path": "bf16 observation written by the step kernel, GEMMs with bias+ReLU "
                             "epilogues, fused output layer" if runner.fused else
                             "actor on the f64 observation (cast + eager layers)",
              "timed": "actor forward + model step + reward, two steps per CUDA-graph replay"})
    model.close()


# --------------------------------------------------------------------------- #
def leg_long_road(ctx, args):
    from mbrl_traffic_b200 import _lib
    from mbrl_traffic_b200.dist import DomainDecomposedRoad, PeerHalo, engine_stepper
    from mbrl_traffic_b200.engine import Engine
    n_road, halo = args.road_cells, args.halo
    steps = args.road_steps - args.road_steps % halo
    params = dict(dx=50, dt=0.5, rho_max=1, v_max=30, lam=1, tau=9)
    road_np = make_obs(1, n_road, 300)[0]
    road = torch.as_tensor(road_np, device=ctx.dev)

    def time_single():
        eng = Engine("arz", "float64", 1, n_road, device=ctx.local_rank)
        eng.set_params(boundary_conditions="extend_both", **params)
        x = road.reshape(1, -1).clone()
        y = torch.empty_like(x)
        eng.rollout(x.clone(), y, 4)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        ref = eng.rollout(x, y, steps)
        e1.record()
        torch.cuda.synchronize()
        return ref.reshape(-1), e0.elapsed_time(e1) * 1e-3

    ref, t_single = time_single()          # every rank: its own copy of the whole road
    out = {"leg": "long_road",
           "workload": "one ARZ road of %d cells, extend_both ends, float64, %d steps "
                       "(BASELINE.json configs[3])" % (n_road, steps),
           "unit": "cell-updates/s", "n_gpus": ctx.world, "steps": steps,
           "single_gpu_value": n_road * steps / ctx.slowest(t_single),
           "scaling": "strong"}
    if ctx.world == 1:
        out.update(value=n_road * steps / t_single, halo=None,
                   roofline_frac=n_road * 32 / (t_single / steps) / 1e9 / ctx.peak)
        ctx.emit(out)
        return
    d = DomainDecomposedRoad(n_road, ctx.rank, ctx.world, halo, None)
    eng = Engine("arz", "float64", 1, d.n_local, device=ctx.local_rank)
    d.stepper = engine_stepper(eng, _lib.MT_BC_EXTEND, _lib.MT_BC_EXTEND, **params)
    peer = PeerHalo(halo, "float64", ctx.local_rank).connect_group(ctx.rank, ctx.world)
    d.peer = peer
    x = d.begin(road)
    d.advance(x, 2 * halo)                  # warm-up incl. two exchanges
    times = []
    for _ in range(3):
        x = d.begin(road)                   # (consumes the previous run's pending push)
        ctx.barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        x = d.advance(x, steps)
        e1.record()
        ctx.barrier()
        times.append(ctx.slowest(e0.elapsed_time(e1) * 1e-3))
    sec = float(np.median(times))
    rho, v = d.own(x)
    same = bool(torch.equal(rho, ref[d.lo:d.hi]) and
                torch.equal(v, ref[n_road + d.lo:n_road + d.hi]))
    out.update(value=n_road * steps / sec, halo=halo, seconds=sec,
               identical_to_single_gpu=ctx.all_ok(same),
               peer_halo_timed_out=not ctx.all_ok(not peer.timed_out()),
               speedup_vs_one_gpu=(n_road * steps / sec) / out["single_gpu_value"],
               exchange="ghost cells stored straight into the neighbours' peer-mapped "
                        "mailboxes (CUDA IPC over NVLink), one exchange per %d steps" % halo)
    ctx.emit(out)
    peer.close()


LEGS = {"nonlocal": leg_nonlocal, "rollout": leg_rollout, "long_road": leg_long_road}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--legs", default="nonlocal,rollout,long_road")
    ap.add_argument("--nonlocal-envs", type=int, default=16384)
    ap.add_argument("--nonlocal-steps", type=int, default=40)
    ap.add_argument("--rollout-envs", type=int, default=4096)
    ap.add_argument("--horizon", type=int, default=500)
    ap.add_argument("--rollout-plain", action="store_true",
                    help="rollout leg without the fused policy path (A/B)")
    ap.add_argument("--road-cells", type=int, default=10_000_000)
    ap.add_argument("--road-steps", type=int, default=256)
    ap.add_argument("--halo", type=int, default=16)
    args = ap.parse_args()
    ctx = Ctx()
    for name in args.legs.split(","):
        t0 = time.time()
        try:
            LEGS[name](ctx, args)
        except Exception as exc:          # report and go on with the next leg
            ctx.emit({"leg": name, "error": repr(exc)[:400]})
            if ctx.dist is not None:      # ranks may have diverged: stop here
                break
        finally:
            if ctx.rank == 0:
                sys.stderr.write("[bench_legs] %s took %.1f s\n" % (name, time.time() - t0))
    if ctx.dist is not None:
        try:
            ctx.dist.barrier()
            ctx.dist.destroy_process_group()
        except Exception:
            pass


if __name__ == "__main__":
    main()
