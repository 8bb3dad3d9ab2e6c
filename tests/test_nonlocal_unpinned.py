"""Non-local look-ahead model: PARITY UNPINNED (the reference names the model,
utils/train.py:685-689, but holds no code for it).  What stands in for
reference vectors here:

* an independent scalar re-implementation of both schemes (pure Python floats,
  cell-major loops, ghost cells resolved by index arithmetic instead of
  padding) that the NumPy oracle must match bit for bit;
* grid convergence on a smooth ring profile (errors against the finest grid
  shrink at first order);
* for lam = 1 the speed law is affine, so the mean of downstream SPEEDS
  (Friedrich-Kolb-Goettlich, Godunov-type scheme) and the speed of the mean
  downstream DENSITY (Blandin-Goatin, Lax-Friedrichs scheme) are the same
  model: the two schemes must approach each other under refinement;
* the CUDA kernels (pipelined and scalar) against the oracle, bit for bit.
"""
from fractions import Fraction

import numpy as np
import pytest

from oracle import np_oracle as o


def _fma(a, b, c):
    """Exact fused multiply-add through rational arithmetic (float() of a
    Fraction rounds to nearest even)."""
    return float(Fraction(a) * Fraction(b) + Fraction(c))

BCS = ["loop", "extend_both", {"constant_both": (0.3, 0.2)}, {"inflow_outflow": 0.3}]


# --------------------------------------------------------------------------- #
# independent scalar restatement                                              #
# --------------------------------------------------------------------------- #
def _rho_at(row, c, bc):
    n = len(row)
    if 0 <= c < n:
        return row[c]
    if bc == "loop":
        return row[c % n]
    if bc == "extend_both":
        return row[0] if c < 0 else row[n - 1]
    kind = list(bc.keys())[0]
    if kind == "constant_both":
        return bc[kind][0] if c < 0 else bc[kind][1]
    if kind == "inflow_outflow":
        return bc[kind] if c < 0 else row[n - 1]
    raise ValueError(bc)


def _v(r, v_max, rho_max, lam):
    x = 1.0 - (r / rho_max)
    p = x if lam == 1 else (x * x if lam == 2 else x ** lam)
    return v_max * p


def scalar_nonlocal(row, gamma, dx, dt, rho_max, v_max, lam, bc, scheme):
    """One step of one road, one cell at a time, Python floats only."""
    row = [float(x) for x in row]
    gamma = [float(g) for g in gamma]
    n = len(row)
    step = dt / dx
    out = [0.0] * n
    for j in range(n):
        if scheme == "godunov":
            face = []
            for c in (j - 1, j):                    # F_{c+1/2} = rho_c W_c
                w = 0.0
                for k, g in enumerate(gamma):
                    w = _fma(g, _v(_rho_at(row, c + k + 1, bc), v_max, rho_max, lam), w)
                face.append(_rho_at(row, c, bc) * w)
            out[j] = row[j] - step * (face[1] - face[0])
        else:
            cell, gq = [], []
            for c in (j - 1, j, j + 1):
                r = 0.0
                for k, g in enumerate(gamma):
                    r = _fma(g, _rho_at(row, c + k, bc), r)
                cell.append(_rho_at(row, c, bc))
                gq.append(cell[-1] * _v(r, v_max, rho_max, lam))
            ha = 0.5 * v_max
            f_lo = 0.5 * (gq[0] + gq[1]) + ha * (cell[0] - cell[1])
            f_hi = 0.5 * (gq[1] + gq[2]) + ha * (cell[1] - cell[2])
            out[j] = row[j] - step * (f_hi - f_lo)
    return np.array(out + [_v(r, v_max, rho_max, lam) for r in out])


def _oracle(scheme):
    return o.nonlocal_step if scheme == "godunov" else o.nonlocal_lxf_step


@pytest.mark.parametrize("scheme", ["godunov", "lxf"])
@pytest.mark.parametrize("bci", [0, 1, 2, 3])
@pytest.mark.parametrize("n,eta,kernel,lam,rho_max", [
    (12, 150, "linear", 1, 1.0), (40, 400, "constant", 2, 0.5), (33, 50, "linear", 1, 0.2),
    (20, 1000, "linear", 2, 1.0)])
def test_oracle_equals_independent_scalar_restatement_parity_unpinned(scheme, bci, n, eta, kernel,
                                                                      lam, rho_max):
    rng = np.random.default_rng(100 * n + bci)
    b = 3
    rho = rho_max * rng.uniform(0.05, 0.9, (b, n))
    obs = np.concatenate((rho, np.zeros_like(rho)), axis=1)
    bc = BCS[bci]
    if isinstance(bc, dict):
        bc = {k: (tuple(rho_max * x for x in v) if isinstance(v, tuple) else rho_max * v)
              for k, v in bc.items()}
    g = o.nonlocal_weights(eta, 50, kernel)
    got = _oracle(scheme)(obs, g, dx=50, dt=0.5, rho_max=rho_max, v_max=25.0, lam=lam,
                          boundary_conditions=bc)
    for i in range(b):
        ref = scalar_nonlocal(rho[i], g, 50, 0.5, rho_max, 25.0, lam, bc, scheme)
        assert np.array_equal(got[i], ref), (scheme, bci, i)


def test_oracle_fma_equals_the_c_library_fma():
    """np_oracle.fma (error-free product + round-to-odd) against libm's fma on
    random and adversarial operands (near-cancellation, long mantissas)."""
    import ctypes
    import ctypes.util
    libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    libm.fma.restype = ctypes.c_double
    libm.fma.argtypes = [ctypes.c_double] * 3
    rng = np.random.default_rng(0)
    n = 20000
    a = np.concatenate([rng.uniform(0, 0.2, n), rng.uniform(-1, 1, n),
                        1 + rng.integers(0, 2 ** 20, n) * 2.0 ** -52])
    b = np.concatenate([rng.uniform(0, 30, n), rng.uniform(-1, 1, n),
                        1 + rng.integers(0, 2 ** 20, n) * 2.0 ** -52])
    c = np.concatenate([rng.uniform(0, 30, n), -rng.uniform(-1, 1, n) * rng.uniform(-1, 1, n),
                        -(1 + rng.integers(0, 2 ** 21, n) * 2.0 ** -52)])
    got = o.fma(a, b, c)
    ref = np.array([libm.fma(x, y, z) for x, y, z in zip(a.tolist(), b.tolist(), c.tolist())])
    assert np.array_equal(got, ref)
    assert _fma(0.1, 0.2, -0.02) == libm.fma(0.1, 0.2, -0.02)


def test_lookahead_weights_have_unit_mass_and_do_not_increase():
    for eta, kernel in ((800, "linear"), (200, "constant"), (50, "linear"), (1600, "linear")):
        g = o.nonlocal_weights(eta, 50, kernel)
        assert abs(g.sum() - 1.0) < 1e-14
        assert np.all(np.diff(g) <= 1e-18)


def _ring_solution(scheme, n, t_end=20.0, length=2000.0, eta=200.0, cfl=0.25):
    """Smooth density bump on a periodic road, advanced to t_end."""
    dx = length / n
    dt = cfl * dx / 30.0
    steps = int(round(t_end / dt))
    x = (np.arange(n) + 0.5) * dx
    rho = 0.3 + 0.25 * np.exp(-((x - 0.4 * length) / (0.08 * length)) ** 2)
    obs = np.concatenate((rho, np.zeros(n)))[None]
    g = o.nonlocal_weights(eta, dx, "linear")
    fn = _oracle(scheme)
    for _ in range(steps):
        obs = fn(obs, g, dx=dx, dt=dt, rho_max=1, v_max=30, lam=1, boundary_conditions="loop")
    return obs[0, :n]


def _coarsen(fine, n):
    return fine.reshape(n, -1).mean(axis=1)


@pytest.mark.parametrize("scheme", ["godunov", "lxf"])
def test_grid_convergence_on_a_smooth_ring_parity_unpinned(scheme):
    """L1 distance to the finest grid roughly halves per refinement."""
    sols = {n: _ring_solution(scheme, n) for n in (40, 80, 160, 640)}
    ref = sols[640]
    err = [np.abs(sols[n] - _coarsen(ref, n)).mean() for n in (40, 80, 160)]
    assert err[0] > err[1] > err[2]
    assert err[0] / err[1] > 1.5 and err[1] / err[2] > 1.5, err
    # mass is conserved on the ring, bounds are kept
    for n, s in sols.items():
        assert abs(s.mean() - sols[640].mean()) < 1e-12
        assert s.min() >= 0.0 and s.max() <= 1.0


def test_both_schemes_approach_the_same_limit_for_an_affine_speed_law():
    """lam = 1: V(sum gamma rho) == sum gamma V(rho), one model, two schemes."""
    dist = []
    for n in (40, 160, 640):
        a, b = _ring_solution("godunov", n), _ring_solution("lxf", n)
        dist.append(np.abs(a - b).mean())
    assert dist[0] > dist[1] > dist[2]
    assert dist[2] < 0.25 * dist[0], dist


# --------------------------------------------------------------------------- #
# CUDA kernels against the oracle (LxF scheme; the Godunov-type scheme is in   #
# tests/test_gpu_parity.py::test_nonlocal_*)                                   #
# --------------------------------------------------------------------------- #
torch = None


def _gpu():
    global torch
    import torch as _t
    torch = _t
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (no CPU fallback exists)")


def _make_obs(b, n, seed, dtype=np.float64):
    rng = np.random.default_rng(seed)
    rho = rng.uniform(0.05, 0.9, (b, n))
    v = 30 * (1 - rho) + rng.normal(0, 1, (b, n))
    return np.concatenate((rho, v), axis=1).astype(dtype)


def _model(**over):
    from mbrl_traffic_b200 import NonLocalModel, NONLOCAL_MODEL_PARAMS
    base = dict(NONLOCAL_MODEL_PARAMS)
    base.pop("optimizer_cls")
    extra = {k: over.pop(k) for k in list(over) if k in ("device", "fast_math")}
    base.update(over)
    return NonLocalModel(None, None, None, None, 0, optimizer_cls=None, **base, **extra)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [32, 64, 250, 1000, 1024, 33, 1001, 5000])
@pytest.mark.parametrize("bci", [0, 1, 2, 3])
@pytest.mark.parametrize("eta,kernel", [(800, "linear"), (200, "constant"), (50, "linear")])
def test_lxf_kernels_match_oracle_f64_parity_unpinned(n, bci, eta, kernel):
    _gpu()
    bc = BCS[bci]
    b = 23 if n <= 1024 else 3
    obs = _make_obs(b, n, seed=7 * n + bci)
    m = _model(eta=eta, kernel=kernel, boundary_conditions=bc, scheme="lxf")
    out = m.get_next_obs(torch.as_tensor(obs, device="cuda:0"), None).cpu().numpy()
    ref = o.nonlocal_lxf_step(obs, o.nonlocal_weights(eta, 50, kernel), boundary_conditions=bc)
    assert np.array_equal(out, ref)
    m.close()


@pytest.mark.gpu
def test_lxf_chained_with_viscosity_and_per_road_parameters():
    _gpu()
    from mbrl_traffic_b200.engine import Engine
    b, n = 48, 500
    rng = np.random.default_rng(16)
    rho_max = rng.choice([1.0, 0.5, 0.2], b)
    v_max = rng.uniform(10, 30, b)
    lam = rng.choice([1.0, 2.0], b)
    rho = rho_max[:, None] * rng.uniform(0.05, 0.9, (b, n))
    obs = np.concatenate((rho, np.zeros_like(rho)), axis=1)
    g = o.nonlocal_weights(400, 50, "linear")
    dev = "cuda:0"
    for alpha in (None, 35.0):
        eng = Engine("nonlocal", "float64", b, n, max_n_eta=64)
        eng.set_params(dx=50, dt=0.5, rho_max=torch.as_tensor(rho_max, device=dev),
                       v_max=torch.as_tensor(v_max, device=dev),
                       lam=torch.as_tensor(lam, device=dev), boundary_conditions="loop",
                       gamma=torch.as_tensor(g, device=dev), scheme="lxf", alpha=alpha)
        x = torch.as_tensor(obs, device=dev)
        y = torch.empty_like(x)
        rew = torch.empty(b, dtype=torch.float64, device=dev)
        ref = obs
        for _ in range(10):
            eng.step(x, y, reward=rew)
            x, y = y, x
            ref = o.nonlocal_lxf_step(ref, g, rho_max=rho_max, v_max=v_max, lam=lam,
                                      alpha=alpha)
        assert np.array_equal(x.cpu().numpy(), ref), alpha
        np.testing.assert_allclose(rew.cpu().numpy(), o.mean_speed_reward(ref), rtol=1e-12)
        eng.close()


@pytest.mark.gpu
def test_lxf_ring_conserves_mass_and_keeps_bounds_full_size():
    _gpu()
    b, n = 16384, 1000
    rng = np.random.default_rng(15)
    rho = rng.uniform(0.05, 0.9, (b, n))
    obs = np.concatenate((rho, np.zeros_like(rho)), axis=1)
    m = _model(scheme="lxf")
    assert m.cfl_number() <= 1.0
    out = m.run(torch.as_tensor(obs, device="cuda:0"), None, n_steps=20).cpu().numpy()
    r = out[:, :n]
    assert r.min() >= 0.0 and r.max() <= 1.0
    np.testing.assert_allclose(r.sum(axis=1), rho.sum(axis=1), rtol=1e-12)
    m.close()


@pytest.mark.gpu
def test_lxf_f32_and_direct_kernel():
    _gpu()
    obs32 = _make_obs(128, 1000, seed=23, dtype=np.float32)
    g = o.nonlocal_weights(800, 50, "linear")
    ref = o.nonlocal_lxf_step(obs32.astype(np.float64), g)
    for fast in (False, True):
        m = _model(scheme="lxf", fast_math=fast)
        out = m.get_next_obs(torch.as_tensor(obs32, device="cuda:0"), None).cpu().numpy()
        err = np.abs(out - ref) / np.maximum(np.abs(ref), 1.0)
        assert err.max() <= 2e-6, err.max()
        m.close()
    # the scalar kernel (MT_FLAG_NO_PIPELINE) gives the pipelined kernel's bits
    from mbrl_traffic_b200.engine import Engine
    obs = _make_obs(16, 1000, seed=24)
    outs = []
    for direct in (False, True):
        eng = Engine("nonlocal", "float64", 16, 1000, max_n_eta=64)
        eng.set_params(dx=50, dt=0.5, rho_max=1, v_max=30, lam=1,
                       boundary_conditions={"inflow_outflow": 0.3},
                       gamma=torch.as_tensor(g, device="cuda:0"), scheme="lxf",
                       no_pipeline=direct)
        x = torch.as_tensor(obs, device="cuda:0")
        y = torch.empty_like(x)
        eng.step(x, y)
        outs.append(y.cpu().numpy())
        eng.close()
    assert np.array_equal(outs[0], outs[1])
    assert np.array_equal(outs[0], o.nonlocal_lxf_step(
        obs, g, boundary_conditions={"inflow_outflow": 0.3}))
