"""Roads longer than one tile: the several-steps-per-launch window kernel
(csrc/mt_segfused.cuh) behind mt_rollout, and the ghost-cell exchange folded
into it (mt_halo_advance) -- emulated ranks in one process, and two real
processes that open each other's mailboxes through CUDA IPC on one GPU."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import c_oracle  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PARAMS = dict(dx=50, dt=0.5, rho_max=1, v_max=30, lam=1, tau=9)


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (no CPU fallback exists)")


def make_obs(b, n, seed=0, dtype=np.float64):
    rng = np.random.default_rng(1234 + seed)
    rho = rng.uniform(0.05, 0.9, (b, n))
    v = 30 * (1 - rho) + rng.normal(0, 1.0, (b, n))
    return np.concatenate((rho, v), axis=1).astype(dtype)


def gpu(x):
    return torch.as_tensor(x, device="cuda:0")


BCS = {"lwr": ["extend_both", {"constant_both": (0.3, 0.2)}, {"inflow_outflow": 0.25}],
       "arz": ["extend_both", {"constant_both": ((0.3, 0.1), (0.2, -0.05))},
               {"inflow_outflow": (0.25, 0.02)}]}


@pytest.mark.parametrize("kind", ["lwr", "arz"])
@pytest.mark.parametrize("b,n", [(1, 1028), (3, 2048), (2, 5000), (1, 40000)])
@pytest.mark.parametrize("bci", [0, 1, 2])
@pytest.mark.parametrize("steps", [2, 16, 37])
def test_window_rollout_equals_step_by_step(kind, b, n, bci, steps, monkeypatch):
    """mt_rollout on roads longer than a tile (overlapping windows, up to 16
    steps per launch) against the same steps launched one by one, f64 and f32,
    and (f64, reference boundary modes) against the oracle."""
    _need_gpu()
    from mbrl_traffic_b200.engine import Engine
    bc = BCS[kind][bci]
    for dtype in ("float64", "float32"):
        obs = make_obs(b, n, seed=n + steps, dtype=np.dtype(dtype))
        eng = Engine(kind, dtype, b, n)
        eng.set_params(boundary_conditions=bc, **PARAMS)
        x = gpu(obs)
        fused = eng.rollout(x.clone(), torch.empty_like(x), steps).clone()
        monkeypatch.setenv("MT_NO_FUSE", "1")
        plain = eng.rollout(x.clone(), torch.empty_like(x), steps).clone()
        monkeypatch.delenv("MT_NO_FUSE")
        assert torch.equal(fused, plain), (dtype, kind, b, n, bci, steps)
        if dtype == "float64" and bci < 2:
            ref = obs
            for _ in range(steps):
                ref = c_oracle.step(kind, ref, boundary_conditions=bc, **PARAMS)
            assert np.array_equal(fused.cpu().numpy(), ref)
        eng.close()


def test_window_rollout_with_per_step_speed_limits_and_per_road_parameters():
    _need_gpu()
    from mbrl_traffic_b200.engine import Engine
    b, n, steps = 3, 6000, 21
    obs = make_obs(b, n, seed=5)
    rng = np.random.default_rng(8)
    rho_max = rng.uniform(0.95, 1.3, b)
    lam = np.array([1.0, 2.0, 1.0])
    acts = rng.uniform(22.0, 30.0, (steps, b))
    eng = Engine("arz", "float64", b, n)
    eng.set_params(boundary_conditions="extend_both", dx=50, dt=0.5, rho_max=gpu(rho_max),
                   v_max=30, lam=gpu(lam), tau=9)
    x = gpu(obs)
    got = eng.rollout(x.clone(), torch.empty_like(x), steps, actions=gpu(acts)).cpu().numpy()
    from oracle import np_oracle as o
    ref = obs
    for j in range(steps):
        ref = o.arz_step(ref, boundary_conditions="extend_both", dx=50, dt=0.5,
                         rho_max=rho_max, v_max=acts[j], lam=lam, tau=9,
                         rho_crit=rho_max / 2)
    assert np.array_equal(got, ref)
    eng.close()


def _ranks(kind, name, n, world, halo, road):
    from mbrl_traffic_b200 import _lib
    from mbrl_traffic_b200.dist import DomainDecomposedRoad, PeerHalo, engine_stepper
    from mbrl_traffic_b200.engine import Engine
    halos = PeerHalo.connect_local([PeerHalo(halo, name, 0) for _ in range(world)])
    ranks, local = [], []
    for r in range(world):
        d = DomainDecomposedRoad(n, r, world, halo, None)
        eng = Engine(kind, name, 1, d.n_local)
        d.stepper = engine_stepper(eng, _lib.MT_BC_EXTEND, _lib.MT_BC_EXTEND, **PARAMS)
        d.peer = halos[r]
        ranks.append(d)
        local.append(d.scatter(gpu(road)).clone())
    return ranks, local, halos


@pytest.mark.parametrize("kind", ["lwr", "arz"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("world,halo,n", [(2, 8, 8192), (3, 16, 24000), (4, 32, 40000),
                                          (8, 4, 100000)])
def test_exchange_folded_into_the_step_launches_emulated(kind, dtype, world, halo, n):
    """mt_halo_advance with all ranks in one process (one stream, one launch
    per rank and exchange period, issued in rank order: every flag a launch
    waits for was published by an earlier launch): the decomposed road equals
    the whole road, cell for cell, also after a period shorter than the halo
    and when the run is continued by the separate pull / step path."""
    _need_gpu()
    if os.environ.get("MT_NO_FUSE") == "1":
        pytest.skip("MT_NO_FUSE=1 switches the several-steps-per-launch kernel off")
    from mbrl_traffic_b200.engine import Engine
    name = np.dtype(dtype).name
    periods, tail = 4, 3
    steps = periods * halo + tail
    road = make_obs(1, n, seed=41, dtype=dtype)[0]
    full = Engine(kind, name, 1, n)
    full.set_params(boundary_conditions="extend_both", **PARAMS)
    a = gpu(road).reshape(1, -1).clone()
    ref = full.rollout(a, torch.empty_like(a), steps + 2).cpu().numpy()[0]
    ref_mid = full.rollout(gpu(road).reshape(1, -1).clone(), torch.empty_like(a),
                           steps).cpu().numpy()[0]
    ranks, local, halos = _ranks(kind, name, n, world, halo, road)
    for d, x in zip(ranks, local):          # (derives the parameter table: one launch)
        d.stepper.buffers(x, "halo" if d.has_left else "physical",
                          "halo" if d.has_right else "physical")
    launches0 = [d.stepper.engine.launch_count for d in ranks]
    for _ in range(periods):
        local = [d.advance(x, halo) for d, x in zip(ranks, local)]
    local = [d.advance(x, tail) for d, x in zip(ranks, local)]     # a short last period
    torch.cuda.synchronize()
    # one launch per rank and period, nothing else
    assert all(d.stepper.engine.launch_count - l0 == periods + 1
               for d, l0 in zip(ranks, launches0))
    assert not any(h.timed_out() for h in halos)
    rho = torch.cat([d.own(x)[0] for d, x in zip(ranks, local)]).cpu().numpy()
    v = torch.cat([d.own(x)[1] for d, x in zip(ranks, local)]).cpu().numpy()
    assert np.array_equal(np.concatenate((rho, v)), ref_mid)
    # continue step by step: the pending push is pulled by the separate kernel
    for d in ranks:
        d.fuse_exchange = False
    for _ in range(2):
        local = [d.step(x) for d, x in zip(ranks, local)]
    torch.cuda.synchronize()
    rho = torch.cat([d.own(x)[0] for d, x in zip(ranks, local)]).cpu().numpy()
    v = torch.cat([d.own(x)[1] for d, x in zip(ranks, local)]).cpu().numpy()
    assert np.array_equal(np.concatenate((rho, v)), ref)
    # a second run on the same objects: begin() consumes the push the first run
    # left pending, so the fresh ghosts are used (not the old run's edge cells)
    for d in ranks:
        d.fuse_exchange = True
    local = [d.begin(gpu(road)) for d in ranks]
    local = [d.advance(x, halo) for d, x in zip(ranks, local)]      # leaves a push pending
    local = [d.begin(gpu(road)) for d in ranks]
    for _ in range(periods):          # (one stream: period by period over the ranks)
        local = [d.advance(x, halo) for d, x in zip(ranks, local)]
    local = [d.advance(x, tail) for d, x in zip(ranks, local)]
    torch.cuda.synchronize()
    assert not any(h.timed_out() for h in halos)
    rho = torch.cat([d.own(x)[0] for d, x in zip(ranks, local)]).cpu().numpy()
    v = torch.cat([d.own(x)[1] for d, x in zip(ranks, local)]).cpu().numpy()
    assert np.array_equal(np.concatenate((rho, v)), ref_mid)
    for h in halos:
        h.close()


def test_halo_advance_argument_checks():
    _need_gpu()
    if os.environ.get("MT_NO_FUSE") == "1":
        pytest.skip("MT_NO_FUSE=1 switches the several-steps-per-launch kernel off")
    from mbrl_traffic_b200 import _lib
    from mbrl_traffic_b200.dist import PeerHalo
    from mbrl_traffic_b200.engine import Engine
    a, b = PeerHalo.connect_local([PeerHalo(8, "float64", 0), PeerHalo(8, "float64", 0)])
    x = torch.zeros(1, 2 * 4000, dtype=torch.float64, device="cuda:0")
    y = torch.empty_like(x)
    eng = Engine("arz", "float64", 1, 4000)
    eng.set_params(boundary_conditions="extend_both", **PARAMS)
    with pytest.raises(ValueError):          # the edge with a neighbour is not a ghost edge
        a.advance(eng, x, y, 8, True)
    eng.set_params(boundary_conditions={"codes": (_lib.MT_BC_EXTEND, _lib.MT_BC_HALO,
                                                  (0.0, 0.0), (0.0, 0.0))}, **PARAMS)
    with pytest.raises(Exception):           # stale ghosts and nothing pushed
        a.advance(eng, x, y, 8, False)
    short = Engine("arz", "float64", 1, 512)  # fits one tile: not the window kernel's case
    short.set_params(boundary_conditions={"codes": (_lib.MT_BC_EXTEND, _lib.MT_BC_HALO,
                                                    (0.0, 0.0), (0.0, 0.0))}, **PARAMS)
    xs = torch.zeros(1, 1024, dtype=torch.float64, device="cuda:0")
    with pytest.raises(ValueError):
        a.advance(short, xs, torch.empty_like(xs), 8, True)
    a.close()
    b.close()


def test_two_processes_exchange_through_cuda_ipc_on_one_gpu(tmp_path):
    """The real inter-process path: two processes (gloo rendezvous on
    127.0.0.1) share GPU 0, open each other's mailboxes with
    mt_halo_handle / mt_halo_connect, and step a decomposed road both ways --
    separate push / pull kernels, and the exchange folded into the step
    launches.  Rank 0 compares the gathered road with one engine stepping the
    whole road."""
    _need_gpu()
    script = os.path.join(ROOT, "tests", "ipc_two_ranks.py")
    port = 29731 + os.getpid() % 200
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK="0",
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, script], env=env, cwd=ROOT,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                      universal_newlines=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            pytest.fail("the two-rank IPC run did not finish in 240 s")
        outs.append(out)
    assert all(p.returncode == 0 for p in procs), "\n".join(o[-2000:] for o in outs)
    assert "IPC_OK" in outs[0], outs[0][-2000:]
