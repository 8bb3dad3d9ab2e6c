#!/usr/bin/env python
"""Headline benchmark: cell-updates/s of the batched ARZ step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--no-extras]

Workload (BASELINE.json configs[1]): ARZ model, 4096 ring roads x 1000 cells
per GPU, float64 (the reference's arithmetic type), reference defaults
(dx=50, dt=0.5, rho_max=1, v_max=30, tau=9, lam=1, "loop" boundaries),
synthetic seeded states.  A "step" is one pass of the hot path over one
batch.  N > 1 (torchrun, one rank per GPU) shards independent roads across
ranks with no data-path collective: weak scaling, value = all ranks' cells.

The per-batch state (65.5 MB) is smaller than the 126 MB L2, so the timed
steps round-robin over 4 independent batches (524 MB working set): every step
reads its input from HBM.  The engine is NOT told that the batches are
independent: it checks by itself, on the host, that a launch's input bytes are
not what the launch before it writes, and only then lets the loads start while
that launch drains (stores always wait).  `dependent_launches` is the same
pool with that check switched off (MT_FLAG_DEPENDENT_INPUT: every launch
waits for the one before it) and `chained` is one batch ping-ponged in place,
where the check finds the dependency and every launch waits as well.

Timing: the K steps are captured in CUDA graphs and the K-step region is
timed REPS (7) times, each repetition bracketed by a barrier + synchronize and
CUDA events on the launching stream, max over ranks; `value` is the MEDIAN
repetition (all repetitions are in `rep_ms_per_step`).  "sustained" repeats
the same launches for ~0.5 s, where the board reaches its power cap.

Beside the headline (same JSON line): `e2e` -- the public
ARZModel.get_next_obs call on pinned host arrays, copies inside the timed
region, with the plain-copy PCIe ceiling measured in the same run;
`e2e_rollout` -- host -> run(n_steps=1000) -> host; `legs` -- BASELINE.json's
other configurations (non-local model, policy-driven rollout, one 10 M-cell
road), measured by tools/bench_legs.py in a child process.

One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import sys
import hashlib
import subprocess
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
REPS = 7                           # repetitions of the timed K-step region (median reported)
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


def csrc_hash():
    """sha256 over the sources the headline kernel is compiled from
    (mt::step_pipe_kernel<double, ARZ>: mt_pipe.cuh and the physics headers)."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "mbrl_traffic_b200", "csrc")
    for name in ("mt_cell.cuh", "mt_math.cuh", "mt_pipe.cuh"):
        with open(os.path.join(d, name), "rb") as f:
            h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()[:16]


def recorded_traffic():
    """DRAM bytes per launch of the step kernel from the committed ncu capture
    (profiles/traffic.json).  The record names the hash of the sources it was
    captured from: if the kernels have changed since, the number is stale and
    None is reported instead."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            rec = json.load(f)
        if rec.get("csrc_sha16") != csrc_hash():
            return None, "profiles/traffic.json was captured from other sources (%s): stale" \
                % rec.get("csrc_sha16")
        return rec.get("arz_f64_4096x1000_bytes_per_launch"), rec.get("source")
    except Exception as exc:
        return None, repr(exc)


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

def pin_rank_to_gpu_cpus(local_rank):
    """Bind this rank to the CPUs NVML lists as local to its GPU (page-locked
    buffers allocated afterwards are first-touched there).  Returns what was
    done, for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64)
                if (word >> b) & 1 and 64 * w + b < n_cpu]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"gpu_local_cpus": len(cpus), "bound_to": len(allowed)}
    except Exception as exc:
        return {"error": repr(exc)[:120]}


def run_legs(args, world):
    """BASELINE.json's other configurations, in a child process (own CUDA
    contexts, own process group): a fault there leaves the headline alone."""
    script = os.path.join(ROOT, "tools", "bench_legs.py")
    env = {k: v for k, v in os.environ.items()
           if not (k.startswith("TORCHELASTIC") or k in (
               "RANK", "LOCAL_RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE", "GROUP_RANK",
               "ROLE_RANK", "ROLE_NAME", "ROLE_WORLD_SIZE", "GROUP_WORLD_SIZE",
               "MASTER_ADDR", "MASTER_PORT", "OMP_NUM_THREADS"))}
    if world == 1:
        cmd = [sys.executable, script]
    else:
        port = int(os.environ.get("MASTER_PORT", "29500")) + 37
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
               "--master-port", str(port), script]
    t0 = time.time()
    legs = {}
    try:
        proc = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                              universal_newlines=True, timeout=args.legs_timeout)
        for ln in proc.stdout.splitlines():
            ln = ln.strip()
            if ln.startswith("{") and '"leg"' in ln:
                try:
                    d = json.loads(ln)
                    legs[d.pop("leg")] = d
                except ValueError:
                    pass
        if proc.returncode != 0:
            legs["_child"] = {"rc": proc.returncode, "stderr_tail": proc.stderr[-400:]}
    except subprocess.TimeoutExpired as exc:
        legs["_child"] = {"error": "timed out after %d s" % args.legs_timeout,
                          "stdout_tail": (exc.stdout or "")[-400:] if isinstance(exc.stdout, str)
                          else ""}
    except Exception as exc:
        legs["_child"] = {"error": repr(exc)[:300]}
    legs["_seconds"] = round(time.time() - t0, 1)
    return legs


def run_ours(args, rank, local_rank, world):
    import torch
    from mbrl_traffic_b200 import ARZModel, ARZ_MODEL_PARAMS
    from mbrl_traffic_b200.engine import Engine, PinnedArray

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    affinity = pin_rank_to_gpu_cpus(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    host_pg = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        host_pg = dist.new_group(backend="gloo")      # host-side waits (no GPU spin)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def slowest(values):
        if world == 1:
            return [float(v) for v in values]
        t = torch.tensor(list(values), dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

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
        """graphs: list of (graph, repeats); returns seconds on the device,
        max over ranks; barrier + synchronize on both sides.  After the
        barrier the GPU has been idle: one untimed replay of the same graph
        goes first on the stream (its launches are NOT between the events), so
        that the K timed steps run at the clocks and cache state of a GPU
        that is already stepping, whatever K is."""
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        barrier()
        with torch.cuda.stream(stream):
            graphs[0][0].replay()                 # untimed pre-roll
            e0.record(stream)
            for g, reps in graphs:
                for _ in range(reps):
                    g.replay()
            e1.record(stream)
        barrier()
        return slowest([e0.elapsed_time(e1) * 1e-3])[0]

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

    # ---- the timed region: exactly K steps, REPS times; the median counts ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_end = time.time() + 0.25
    rep_s = [timed(plan) for _ in range(REPS)]
    while time.time() < t_end:                   # short K: keep the same load up
        for g, reps in plan:                     # (untimed) so NVML gets samples
            g.replay()
        torch.cuda.synchronize()
    clocks = sampler.stop()
    seconds = float(np.median(rep_s))

    # ---- sustained: the same graph for ~0.5 s (the board reaches its power cap)
    time.sleep(0.5)
    sus_sampler = ClockSampler(local_rank)
    sus_sampler.start()
    sus_reps = max(1, 20000 // gk)
    sus_s = timed([(plan[0][0], sus_reps)])
    sus_clocks = sus_sampler.stop()
    sus_steps = sus_reps * gk

    # ---- dependent: the same pool, every launch waits for the one before it ---
    eng.set_params(dependent_input=True, **PARAMS)
    with torch.cuda.stream(stream):
        launch_streaming(N_SETS)                  # (eager: consumes the parameter event)
    torch.cuda.synchronize()
    dep_plan = [(capture(lambda k: launch_streaming(k), gk), max(1, full))]
    dep_plan[0][0].replay()
    dep_s = float(np.median([timed(dep_plan) for _ in range(3)]))
    dep_steps = gk * max(1, full)
    eng.set_params(**PARAMS)
    with torch.cuda.stream(stream):
        launch_streaming(N_SETS)
    torch.cuda.synchronize()

    # ---- chained: one batch ping-ponged in place (what a rollout sees) --------
    chain_k = gk + (gk & 1)
    chain_plan = [(capture(launch_chained, chain_k), max(1, full))]
    chain_plan[0][0].replay()
    chain_s = float(np.median([timed(chain_plan) for _ in range(3)]))
    chain_steps = chain_k * max(1, full)
    ins[0].copy_(torch.as_tensor(make_obs(rank * N_SETS), device=dev))
    torch.cuda.synchronize()

    # ---- e2e: the public call on HOST arrays, copies inside the timed region --
    p = dict(ARZ_MODEL_PARAMS)
    p.pop("optimizer_cls")
    model = ARZModel(None, None, None, None, 0, optimizer_cls=None,
                     device=local_rank, **p)
    obs_bytes = N_ENVS * 2 * N_CELLS * 8
    pin_in = PinnedArray((N_ENVS, 2 * N_CELLS), np.float64)
    pin_in.array[...] = make_obs(rank * N_SETS)
    e2e_steps = max(3, min(K, 20))
    for _ in range(3):
        out = model.get_next_obs(pin_in.array, None)   # NumPy in -> new NumPy array out
    del out
    e2e_reps = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            out = model.get_next_obs(pin_in.array, None)
        torch.cuda.synchronize()
        e2e_reps.append(time.perf_counter() - t0)
        barrier()
    e2e_s = float(np.median(slowest(e2e_reps)))
    e2e_checksum = float(out[0, :4].sum())       # the result is really on the host
    del out

    # the ceiling of that call: nothing but the two copies, both directions at
    # once on two streams, same bytes, same pinned buffers, all ranks together
    pin_out = PinnedArray((N_ENVS, 2 * N_CELLS), np.float64)
    h_in, h_out = torch.from_numpy(pin_in.array), torch.from_numpy(pin_out.array)
    d_in, d_out = torch.empty_like(ins[0]), torch.empty_like(ins[0])
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    copy_reps = []
    for rep in range(4):
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        if rep:
            copy_reps.append(time.perf_counter() - t0)
        barrier()
    copy_s = float(np.median(slowest(copy_reps)))

    # ---- e2e_rollout: host -> 1000 steps on the GPU -> host --------------------
    roll_steps = 1000
    out = model.run(pin_in.array, None, n_steps=roll_steps)
    del out
    barrier()
    t0 = time.perf_counter()
    out = model.run(pin_in.array, None, n_steps=roll_steps)
    torch.cuda.synchronize()
    roll_s = slowest([time.perf_counter() - t0])[0]
    roll_finite = bool(np.isfinite(out[:64]).all())
    del out

    cells = N_ENVS * N_CELLS
    value = world * cells * K / seconds
    per_launch_s = seconds / K
    peak, peak_src = measured_peak()
    achieved = cells * BYTES_PER_CELL_UPDATE / per_launch_s / 1e9
    traffic, traffic_src = recorded_traffic()
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
            "launch": "CUDA graph of %d step kernels, replayed; no hint from the caller: the "
                      "engine checks on the host that a launch's input is not the previous "
                      "launch's output before it lets the loads start early" % gk,
            "timing": "median of %d repetitions of the K-step region, each between "
                      "barriers (one untimed replay of the graph first, then the events "
                      "around exactly K steps), CUDA events, max over ranks" % REPS,
        },
        "reps": REPS,
        "rep_ms_per_step": [round(x / K * 1e3, 6) for x in rep_s],
        "gpu_launches": int(K) if captured == gk + rem else int(captured),
        "clocks": clocks,
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": peak_src,
            "peak_spec": 8000.0, "frac_of_spec": achieved / 8000.0,
            "kernel": "mt::step_pipe_kernel<double, ARZ>",
            "algorithmic_bytes_per_cell_update": BYTES_PER_CELL_UPDATE,
            "best_rep_frac": cells * BYTES_PER_CELL_UPDATE / (min(rep_s) / K) / 1e9 / peak,
        },
        "sustained": {"value": world * cells * sus_steps / sus_s, "unit": "cell-updates/s",
                      "steps": sus_steps, "ms_per_step": sus_s / sus_steps * 1e3,
                      "roofline_frac": cells * BYTES_PER_CELL_UPDATE / (sus_s / sus_steps) / 1e9 / peak,
                      "clocks": sus_clocks},
        "dependent_launches": {"value": world * cells * dep_steps / dep_s,
                               "unit": "cell-updates/s", "ms_per_step": dep_s / dep_steps * 1e3,
                               "roofline_frac": cells * BYTES_PER_CELL_UPDATE / (dep_s / dep_steps) / 1e9 / peak,
                               "note": "the same pool with MT_FLAG_DEPENDENT_INPUT: every launch "
                                       "waits for the one before it"},
        "chained": {"value": world * cells * chain_steps / chain_s,
                    "unit": "cell-updates/s", "ms_per_step": chain_s / chain_steps * 1e3,
                    "note": "one batch ping-ponged in place (rollout pattern; "
                            "state partly L2-resident)"},
        "e2e": {"value": world * cells * e2e_steps / e2e_s,
                "unit": "cell-updates/s", "h2d_bytes_per_step": obs_bytes,
                "d2h_bytes_per_step": obs_bytes, "steps": e2e_steps,
                "ms_per_step": e2e_s / e2e_steps * 1e3,
                "api": "ARZModel.get_next_obs(numpy_array, None) -> new numpy array "
                       "(input page-locked; result on a recycled page-locked block)",
                "result_checksum": e2e_checksum,
                "pcie_ceiling_ms_per_step": copy_s / e2e_steps * 1e3,
                "pcie_ceiling_GBps_each_way": obs_bytes / (copy_s / e2e_steps) / 1e9,
                "pcie_frac": copy_s / e2e_s,
                "pcie_note": "ceiling = the same bytes as two plain concurrent pinned copies "
                             "(H2D + D2H on two streams), all %d ranks at once" % world,
                "cpu_affinity": affinity},
        "e2e_rollout": {"value": world * cells * roll_steps / roll_s, "unit": "cell-updates/s",
                        "n_steps": roll_steps, "seconds": roll_s, "finite": roll_finite,
                        "api": "ARZModel.run(numpy_array, None, n_steps=1000): host -> 1000 "
                               "steps in one fused launch per chunk -> host "
                               "(the loop of evaluate_model.py:398-403)"},
    }
    pin_in.free()
    pin_out.free()
    model.close()

    # ---- the other BASELINE.json configurations, in a child process ------------
    if not args.no_extras:
        del ins, outs, d_in, d_out
        torch.cuda.empty_cache()
        if rank == 0:
            line["legs"] = run_legs(args, world)
        if host_pg is not None:
            dist.barrier(group=host_pg)           # the other ranks wait on the host

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
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the child process that measures the other configurations")
    ap.add_argument("--legs-timeout", type=int, default=420)
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
