#!/usr/bin/env python
"""Headline benchmark: cell-updates/s of the batched ARZ step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): ARZ model, 4096 ring roads x 1000 cells
per GPU, float64 (the reference's arithmetic type), reference defaults
(dx=50, dt=0.5, rho_max=1, v_max=30, tau=9, lam=1, "loop" boundaries),
synthetic seeded states.  A "step" is one pass of the hot path over one
batch.  N > 1 (torchrun, one rank per GPU) shards independent roads across
ranks with no data-path collective: weak scaling, value = all ranks' cells.

The per-batch state (65.5 MB) is smaller than the 126 MB L2, so the timed
steps round-robin over 4 independent batches (524 MB working set): every step
reads its input from HBM.  The rate on one batch ping-ponged in place (what a
rollout sees, partly L2-resident) is reported separately as "chained".

The default timed region is 2000 steps (~50 ms, burst clocks -- the regime in
which the roofline's measured peak was taken); "sustained" repeats the same
launches for ~0.5 s, where the board reaches its power cap, and is reported
next to it with the clocks seen.

One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_ENVS = 4096
N_CELLS = 1000
N_SETS = 4
DTYPE = "float64"
BYTES_PER_CELL_UPDATE = 4 * 8      # ARZ: read + write of rho and v
GRAPH_STEPS = 200                  # kernel launches captured per CUDA graph
PARAMS = dict(dx=50, dt=0.5, rho_max=1, v_max=30, tau=9, lam=1,
              boundary_conditions="loop")
WORKLOAD = ("ARZ, %d ring envs x %d cells per GPU, float64, 'loop' boundaries "
            "(BASELINE.json configs[1])" % (N_ENVS, N_CELLS))


def make_obs(seed, n_envs=N_ENVS, n_cells=N_CELLS):
    """rho ~ U(0.05, 0.9) rho_max, v = V(rho) + N(0, 1)  (SURVEY.md 8d)."""
    rng = np.random.default_rng(1234 + seed)
    rho = rng.uniform(0.05, 0.9, (n_envs, n_cells))
    v = 30.0 * (1.0 - rho) + rng.normal(0.0, 1.0, (n_envs, n_cells))
    return np.concatenate((rho, v), axis=1)


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic():
    """DRAM bytes per launch of the step kernel from the committed ncu capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get("arz_f64_4096x1000_bytes_per_launch")
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Polls SM clock and throttle reasons through NVML while a region runs."""

    def __init__(self, index):
        super(ClockSampler, self).__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.error = None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {
                "hw_slowdown": pynvml.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonSwThermalSlowdown,
                "sw_power_cap": pynvml.nvmlClocksThrottleReasonSwPowerCap,
            }
            while not self._stop_evt.is_set():
                self.samples.append(
                    pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for name, bit in names.items():
                    if mask & bit:
                        self.reasons.add(name)
                time.sleep(0.01)
        except Exception as exc:  # NVML missing: report, do not fail the bench
            self.error = repr(exc)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        out = {"sm_mhz": (float(np.median(self.samples)) if self.samples else None),
               "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
               "samples": len(self.samples)}
        if self.error:
            out["error"] = self.error
        return out


# --------------------------------------------------------------------------- #
# CPU arm: the oracle port, all host threads                                  #
# --------------------------------------------------------------------------- #

def time_cpu_port(steps, warmup, min_seconds=3.0):
    """Times oracle/c_oracle.c (OpenMP, all host threads) on the full
    workload shape; a bounded number of steps."""
    from oracle import c_oracle
    # all the host threads this process may use -- torchrun exports
    # OMP_NUM_THREADS=1 to its workers, which would silently make this a
    # one-core baseline
    try:
        threads = len(os.sched_getaffinity(0))
    except AttributeError:
        threads = os.cpu_count() or 1
    threads = max(threads, c_oracle.max_threads())
    obs = make_obs(0)
    out = np.empty_like(obs)
    kw = dict(PARAMS, n_threads=threads)
    for _ in range(max(1, warmup)):
        c_oracle.step("arz", obs, out=out, **kw)
        obs, out = out, obs
    done, t0 = 0, time.perf_counter()
    while True:
        c_oracle.step("arz", obs, out=out, **kw)
        obs, out = out, obs
        done += 1
        dt = time.perf_counter() - t0
        if done >= steps and (dt >= min_seconds or done >= 50 * max(1, steps)):
            break
        if dt > 60.0:
            break
    per_step = dt / done
    value = N_ENVS * N_CELLS / per_step
    return dict(value=value, unit="cell-updates/s", cores=threads, kind="port",
                sample="%d steps of the full %dx%d batch with oracle/c_oracle.c "
                       "(closed-form relaxation, OpenMP over roads); the "
                       "reference's own NumPy+fsolve step runs at ~1.07e3 "
                       "cell-updates/s (BASELINE.md)" % (done, N_ENVS, N_CELLS)), per_step


def run_reference(args, rank, world):
    if rank != 0:
        return
    base, per_step = time_cpu_port(max(1, min(args.steps, 20)), min(args.warmup, 3))
    line = {
        "impl": "reference",
        "metric": "cell_updates_per_s", "value": base["value"],
        "unit": "cell-updates/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_envs_per_gpu": N_ENVS,
                   "n_cells": N_CELLS, "parallelism": "cpu threads"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "cell-updates/s",
                "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(json.dumps(line))


# --------------------------------------------------------------------------- #
# GPU arm                                                                     #
# --------------------------------------------------------------------------- #

def run_ours(args, rank, local_rank, world):
    import torch
    from mbrl_traffic_b200 import ARZModel, ARZ_MODEL_PARAMS
    from mbrl_traffic_b200.engine import Engine, PinnedArray

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng = Engine("arz", DTYPE, N_ENVS, N_CELLS, device=local_rank)
    eng.set_params(**PARAMS)
    ins = [torch.as_tensor(make_obs(rank * N_SETS + j), device=dev)
           for j in range(N_SETS)]
    outs = [torch.empty_like(x) for x in ins]
    stream = torch.cuda.Stream(device=dev)

    def launch_streaming(k, start=0):
        for i in range(start, start + k):
            j = i % N_SETS
            eng.step(ins[j], outs[j], stream=stream)

    def launch_chained(k):
        a, b = ins[0], outs[0]
        for _ in range(k):
            eng.step(a, b, stream=stream)
            a, b = b, a

    def capture(fn, k):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            fn(k)
        return g

    def timed(graphs):
        """graphs: list of (graph, repeats); returns seconds on the device."""
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        barrier()
        with torch.cuda.stream(stream):
            e0.record(stream)
            for g, reps in graphs:
                for _ in range(reps):
                    g.replay()
            e1.record(stream)
        barrier()
        return e0.elapsed_time(e1) * 1e-3

    K, W = args.steps, max(args.warmup, 3)
    with torch.cuda.stream(stream):
        launch_streaming(W)                       # warm-up (also first-use set-up)
    torch.cuda.synchronize()
    launches0 = eng.launch_count
    gk = min(K, GRAPH_STEPS)
    full, rem = divmod(K, gk)
    plan = [(capture(lambda k: launch_streaming(k), gk), full)]
    if rem:
        plan.append((capture(lambda k: launch_streaming(k), rem), 1))
    captured = eng.launch_count - launches0     # kernels recorded into the graphs
    for g, _ in plan:                            # graph warm-up (untimed)
        g.replay()
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    t_end = time.time() + 0.25
    seconds = timed(plan)                        # the timed region: exactly K steps
    while time.time() < t_end:                   # short K: keep the same load up
        for g, reps in plan:                     # (untimed) so NVML gets samples
            g.replay()
        torch.cuda.synchronize()
    clocks = sampler.stop()

    # independent batches: the timed launches DO read batches that the previous
    # launch did not write, which MT_FLAG_INDEPENDENT_INPUT lets the engine use
    # (the next launch fetches its input while the previous one drains).  Not
    # the headline: a rollout's input is the previous step's output.
    time.sleep(1.5)      # let the board's power averaging recover: same (burst) state as the headline
    eng.set_params(independent_input=True, **PARAMS)
    with torch.cuda.stream(stream):
        launch_streaming(N_SETS)
    torch.cuda.synchronize()
    ind_graph = capture(lambda k: launch_streaming(k), gk)
    ind_graph.replay()
    ind_reps = max(1, min(K, 2000) // gk)
    ind_s = timed([(ind_graph, ind_reps)])
    ind_steps = ind_reps * gk
    eng.set_params(**PARAMS)
    with torch.cuda.stream(stream):
        launch_streaming(1)                       # (consumes the parameter-ready event
    torch.cuda.synchronize()                      #  outside of any graph capture)

    # sustained: the same graph for ~0.5 s, long enough for the board to reach
    # its power cap (the headline region is short, like the burst measurement
    # behind MEASURED_PEAKS.json); reported beside the headline, with clocks
    sus_sampler = ClockSampler(local_rank)
    sus_sampler.start()
    sus_reps = max(1, 20000 // gk)
    sus_s = timed([(plan[0][0], sus_reps)])
    sus_clocks = sus_sampler.stop()
    sus_steps = sus_reps * gk

    # chained: one batch ping-ponged in place
    chain_plan = [(capture(launch_chained, gk + (gk & 1)), max(1, full))]
    chain_plan[0][0].replay()
    chain_s = timed(chain_plan)
    chain_steps = (gk + (gk & 1)) * max(1, full)
    ins[0].copy_(torch.as_tensor(make_obs(rank * N_SETS), device=dev))

    # e2e: the public get_next_obs call on pinned HOST buffers, copies included
    p = dict(ARZ_MODEL_PARAMS)
    p.pop("optimizer_cls")
    model = ARZModel(None, None, None, None, 0, optimizer_cls=None,
                     device=local_rank, **p)
    pin_in = PinnedArray((N_ENVS, 2 * N_CELLS), np.float64)
    pin_in.array[...] = make_obs(rank * N_SETS)
    host_eng = model._get_engine(N_ENVS, N_CELLS, DTYPE, host=True)
    pin_out = PinnedArray((N_ENVS, 2 * N_CELLS), np.float64)
    e2e_steps = max(3, min(K, 20))
    for _ in range(2):
        host_eng.step_host(pin_in.array, pin_out.array)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_eng.step_host(pin_in.array, pin_out.array)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    obs_bytes = N_ENVS * 2 * N_CELLS * 8

    if world > 1:
        t = torch.tensor([seconds, chain_s, e2e_s, sus_s, ind_s], dtype=torch.float64,
                         device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        seconds, chain_s, e2e_s, sus_s, ind_s = [float(x) for x in t.tolist()]

    cells = N_ENVS * N_CELLS
    value = world * cells * K / seconds
    per_launch_s = seconds / K
    peak, peak_src = measured_peak()
    achieved = cells * BYTES_PER_CELL_UPDATE / per_launch_s / 1e9
    line = {
        "metric": "cell_updates_per_s", "value": value,
        "unit": "cell-updates/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": per_launch_s * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": WORKLOAD, "n_envs_per_gpu": N_ENVS, "n_cells": N_CELLS,
            "parallelism": "env-sharded x%d, no data-path collective" % world,
            "l2": "inputs larger than L2: %d independent batches round-robin, "
                  "%.0f MB working set per GPU" % (N_SETS, 2 * N_SETS * obs_bytes / 1e6),
            "launch": "CUDA graph of %d step kernels, replayed" % gk,
        },
        "gpu_launches": int(K) if captured == gk + rem else int(captured),
        "clocks": clocks,
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": recorded_traffic(),
            "peak_source": peak_src,
            "peak_spec": 8000.0, "frac_of_spec": achieved / 8000.0,
            "kernel": "mt::step_pipe_kernel<double, ARZ>",
            "algorithmic_bytes_per_cell_update": BYTES_PER_CELL_UPDATE,
        },
        "sustained": {"value": world * cells * sus_steps / sus_s, "unit": "cell-updates/s",
                      "steps": sus_steps, "ms_per_step": sus_s / sus_steps * 1e3,
                      "roofline_frac": cells * BYTES_PER_CELL_UPDATE / (sus_s / sus_steps) / 1e9 / peak,
                      "clocks": sus_clocks},
        "independent_batches": {
            "value": world * cells * ind_steps / ind_s, "unit": "cell-updates/s",
            "ms_per_step": ind_s / ind_steps * 1e3,
            "roofline_frac": cells * BYTES_PER_CELL_UPDATE / (ind_s / ind_steps) / 1e9 / peak,
            "note": "same launches with MT_FLAG_INDEPENDENT_INPUT: a step's loads start "
                    "while the previous launch drains (valid because consecutive timed "
                    "steps work on different batches; not valid for chained steps)"},
        "chained": {"value": world * cells * chain_steps / chain_s,
                    "unit": "cell-updates/s",
                    "note": "one batch ping-ponged in place (rollout pattern; "
                            "state partly L2-resident)"},
        "e2e": {"value": world * cells * e2e_steps / e2e_s,
                "unit": "cell-updates/s", "h2d_bytes_per_step": obs_bytes,
                "d2h_bytes_per_step": obs_bytes, "steps": e2e_steps,
                "api": "ARZModel.get_next_obs on pinned host arrays "
                       "(mt_step_host)"},
    }
    if rank == 0:
        base, _ = time_cpu_port(3, 1) if world == 1 else (None, None)
        if base is not None:
            # the same leg doubles as the run's parity check: one step of the
            # full batch, CUDA path vs the CPU port, compared bit for bit
            from oracle import c_oracle
            obs0 = make_obs(0)
            x = torch.as_tensor(obs0, device=dev)
            y = torch.empty_like(x)
            eng.step(x, y)
            torch.cuda.synchronize()
            ref = c_oracle.step("arz", obs0, **PARAMS)
            got = y.cpu().numpy()
            base["parity"] = {"what": "1 step of the full batch, CUDA vs CPU port",
                              "bit_equal": bool(np.array_equal(got, ref)),
                              "max_abs_diff": float(np.max(np.abs(got - ref)))}
            line["cpu_baseline"] = base
        emit(json.dumps(line))
    pin_in.free()
    pin_out.free()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_OUT_FD = None


def emit(text):
    """The one JSON line goes to the process's real stdout; everything else
    written to fd 1 meanwhile -- NCCL prints its version there when
    NCCL_DEBUG=VERSION -- was sent to stderr by main()."""
    sys.stdout.flush()
    data = (text + "\n").encode()
    if _OUT_FD is None:
        os.write(1, data)
    else:
        os.write(_OUT_FD, data)


def main():
    global _OUT_FD
    sys.stdout.flush()
    _OUT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 and args.gpus > 1:
        raise SystemExit(
            "launch multi-GPU runs with: python -m torch.distributed.run "
            "--nnodes=1 --nproc-per-node %d --master-addr 127.0.0.1 "
            "--master-port 29500 bench.py --gpus %d ..." % (args.gpus, args.gpus))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
