"""Batched NumPy restatement of the reference finite-volume step (oracle).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Every function follows a
reference function and keeps its exact operation order so that fp64 results
are bit-comparable; the only deliberate departure is the ARZ relaxation,
where the reference's ``scipy.optimize.fsolve`` call (arz.py:462-463) is
replaced by the closed-form root of the same affine, diagonal system
(documented at ``arz_step``).

Layout: ``obs`` is ``(B, 2N)`` (or ``(2N,)``), each row ``[rho_0..rho_{N-1} |
v_0..v_{N-1}]`` exactly like the reference (lwr.py:197,204; arz.py:210-211,221).
Model parameters may be Python scalars or per-row arrays of shape ``(B,)``.

Parity status: LWR / ARZ pinned by the golden files generated from the
reference (``oracle/make_golden.py``).  Non-local model: **parity unpinned**
(no reference implementation exists); restated from the paper cited at
lwr.py:209-211.
"""
import numpy as np

ZERO_DENSITY_PATCH = 0.0000000001  # arz.py:275


# --------------------------------------------------------------------------- #
# helpers                                                                     #
# --------------------------------------------------------------------------- #

def _col(p, dtype=None):
    """Turn a scalar or (B,) parameter into something that broadcasts on (B,N).

    Python scalars are returned unchanged so that NumPy's weak-scalar rules
    (and its scalar-exponent fast paths) behave exactly as in the reference,
    where every parameter is a Python number.
    """
    if np.ndim(p) == 0:
        return p
    a = np.asarray(p, dtype=dtype)
    return a.reshape(-1, 1)


def _pow(x, lam):
    """``x ** lam`` as the reference evaluates it (lwr.py:316, arz.py:256).

    With a scalar exponent NumPy takes exact fast paths for 1, 2 and 0.5
    (identity, square, sqrt).  With per-row exponents we apply the scalar
    form row group by row group so the result is the same as B independent
    reference models.
    """
    if np.ndim(lam) == 0:
        return x ** lam
    lam = np.asarray(lam).reshape(-1)
    x = np.asarray(x)
    if x.ndim < 2:
        x = x.reshape(-1, 1)
    if x.shape[0] != lam.shape[0]:
        x = np.broadcast_to(x, (lam.shape[0], x.shape[1]))
    out = np.empty(x.shape, dtype=x.dtype)
    for val in np.unique(lam):
        rows = lam == val
        out[rows] = x[rows] ** float(val)
    return out


def v_eq(rho, v_max, rho_max, lam):
    """Greenshields equilibrium speed (lwr.py:299-316, arz.py:240-256)."""
    return _col(v_max) * _pow(1 - (rho / _col(rho_max)), lam)


def _split(obs):
    obs = np.asarray(obs)
    squeeze = obs.ndim == 1
    if squeeze:
        obs = obs[None, :]
    n = obs.shape[1] // 2
    return obs[:, :n], obs[:, n:2 * n], squeeze


def _shift_left_dup(a):
    """``np.append(a[1:], a[-1])`` per row (lwr.py:294, arz.py:405)."""
    return np.concatenate((a[:, 1:], a[:, -1:]), axis=1)


def _shift_right_dup(a):
    """``np.insert(a[:-1], 0, a[0])`` per row (lwr.py:231, arz.py:414-415)."""
    return np.concatenate((a[:, :1], a[:, :-1]), axis=1)


def _bc_kind(bc):
    if bc == "loop":
        return "loop"
    if bc == "extend_both":
        return "extend_both"
    if isinstance(bc, dict) and list(bc.keys())[0] == "constant_both":
        return "constant_both"
    return None


# --------------------------------------------------------------------------- #
# LWR (lwr.py)                                                                #
# --------------------------------------------------------------------------- #

def lwr_godunov_flux(rho, v_max, rho_max, lam):
    """Demand/supply Godunov flux (lwr.py:268-297).

    Keeps the reference's per-cell ``q_max = rho_crit * V(rho_i)`` and the
    mask-multiply form of the demand and supply.
    """
    rho_crit = _col(rho_max) / 2
    ve = v_eq(rho, v_max, rho_max, lam)
    q_max = rho_crit * ve
    d = rho * ve * (rho < rho_crit) + q_max * (rho >= rho_crit)
    s = rho * ve * (rho > rho_crit) + q_max * (rho <= rho_crit)
    s = _shift_left_dup(s)
    return np.minimum(d, s)


def lwr_ibvp(rho, dx, dt, v_max, rho_max, lam, boundary_conditions):
    """One Godunov step of the density (lwr.py:206-266)."""
    step = dt / dx
    f = lwr_godunov_flux(rho, v_max, rho_max, lam)
    fm = _shift_right_dup(f)
    rho_tp1 = rho - step * (f - fm)

    kind = _bc_kind(boundary_conditions)
    if kind == "loop":
        # lwr.py:237-244 -- ghost copies of the UPDATED values.
        rho_tp1 = np.concatenate(
            (rho_tp1[:, -1:], rho_tp1[:, 1:-1], rho_tp1[:, -2:-1]), axis=1)
    elif kind == "extend_both":
        pass  # lwr.py:248-249
    elif kind == "constant_both":
        left, right = boundary_conditions["constant_both"]
        b = rho_tp1.shape[0]
        rho_tp1 = np.concatenate(
            (np.full((b, 1), left, dtype=rho_tp1.dtype),
             rho_tp1[:, 1:-1],
             np.full((b, 1), right, dtype=rho_tp1.dtype)), axis=1)
    else:
        raise ValueError("Unknown boundary conditions: {}".format(
            boundary_conditions))
    return rho_tp1


def lwr_step(obs, dx=50, dt=0.5, rho_max=1, v_max=30, lam=1,
             boundary_conditions="loop"):
    """``LWRModel.get_next_obs`` for a batch (lwr.py:192-204).

    The velocity half of the input is ignored, as in the reference.
    """
    rho, _, squeeze = _split(obs)
    rho_tp1 = lwr_ibvp(rho, dx, dt, v_max, rho_max, lam, boundary_conditions)
    v_tp1 = v_eq(rho_tp1, v_max, rho_max, lam)
    out = np.concatenate((rho_tp1, v_tp1), axis=1)
    return out[0] if squeeze else out


# --------------------------------------------------------------------------- #
# ARZ (arz.py)                                                                #
# --------------------------------------------------------------------------- #

def arz_relative_flow(rho, v, v_max, rho_max, lam):
    """``y = (v - V(rho)) * rho`` (arz.py:223-238)."""
    return (v - v_eq(rho, v_max, rho_max, lam)) * rho


def arz_compute_flux(rho, y, v_max, rho_max, lam, rho_crit):
    """Supply/demand flux of (rho, y) (arz.py:368-417)."""
    rho_crit = _col(rho_crit)
    ve = v_eq(rho, v_max, rho_max, lam)
    q_max = (rho_crit * v_eq(rho_crit, v_max, rho_max, lam)) + y
    d = (rho * ve + y) * (rho < rho_crit) + q_max * (rho >= rho_crit)
    s = (rho * ve + y) * (rho > rho_crit) + q_max * (rho <= rho_crit)
    s = _shift_left_dup(s)
    q = np.minimum(d, s)
    with np.errstate(divide="ignore", invalid="ignore"):
        p = q * (y / rho)
    qm = _shift_right_dup(q)
    pm = _shift_right_dup(p)
    return q, qm, p, pm


def arz_step(obs, dx=50, dt=0.5, rho_max=1, v_max=30, tau=9, lam=1,
             boundary_conditions="loop", rho_crit=None, fsolve_abort=False):
    """``ARZModel.get_next_obs`` for a batch (arz.py:205-221).

    Relaxation.  The reference solves ``F(z) = z + (dt/tau) *
    (rho' * u(rho', z)) - rhs = 0`` with fsolve (arz.py:462-463, 469-492),
    where ``rhs = y + step*(Pm - P) + (dt/tau)*rho'*V(rho')`` (arz.py:460-461)
    and ``rho' * u(rho', z) = z + rho' * V(rho')`` (arz.py:277).  The
    ``rho' V(rho')`` terms cancel, F is affine and diagonal, and its root is

        y' = (y + step * (Pm - P)) / (1 + dt / tau).

    This closed form is what the oracle (and the CUDA kernel) evaluate.  It
    equals the fsolve output to a few ulp in the velocity half; the density
    half does not depend on the solve and is bit-equal.

    ``rho_crit`` defaults to ``rho_max / 2`` and is a separate argument
    because the reference freezes it at construction (arz.py:199).

    Domain: rho > 0 everywhere.  An exact zero density makes ``P`` NaN here
    (0/0, arz.py:411) and with it ``rhs`` (arz.py:460); every residual fsolve
    then sees is NaN, MINPACK rejects every trial step and hands back its
    initial guess ``x0 = y`` for the WHOLE road (the system is solved jointly,
    arz.py:462-463) -- the density still advances.  ``fsolve_abort=True``
    mimics that: a road whose ``rhs`` holds a non-finite value in ANY of its N
    cells keeps ``y' = y`` [checked against the unmodified reference:
    tests/golden, cases ``arz_zero_*``].
    """
    rho, v, squeeze = _split(obs)
    if rho_crit is None:
        rho_crit = _col(rho_max) / 2
    step = dt / dx

    y = arz_relative_flow(rho, v, v_max, rho_max, lam)
    q, qm, p, pm = arz_compute_flux(rho, y, v_max, rho_max, lam, rho_crit)

    # arz.py:456 and the closed-form root; interior cells only (arz.py:464-465)
    rho_new = rho + (step * (qm - q))
    y_new = (y + (step * (pm - p))) / (1 + dt / _col(tau))
    if fsolve_abort:
        with np.errstate(invalid="ignore", over="ignore"):
            rhs = y + (step * (pm - p)) + (dt / _col(tau)) * (
                rho_new * v_eq(rho_new, v_max, rho_max, lam))      # arz.py:460-461
        aborted = ~np.isfinite(rhs).all(axis=1)
        y_new = np.where(aborted[:, None], y, y_new)
    rho_in = rho_new[:, 1:-1]
    y_in = y_new[:, 1:-1]

    kind = _bc_kind(boundary_conditions)
    b = rho.shape[0]
    if kind == "loop":
        # arz.py:316-328 -- copies of the OLD state.
        rl, yl = rho[:, -1:], y[:, -1:]
        rr, yr = rho[:, -2:-1], y[:, -2:-1]
    elif kind == "extend_both":
        # arz.py:332-344
        rl, yl = rho_in[:, :1], y_in[:, :1]
        rr, yr = rho_in[:, -1:], y_in[:, -1:]
    elif kind == "constant_both":
        # arz.py:347-360 -- second tuple entries are relative flows.
        left, right = boundary_conditions["constant_both"]
        rl = np.full((b, 1), left[0], dtype=rho.dtype)
        yl = np.full((b, 1), left[1], dtype=rho.dtype)
        rr = np.full((b, 1), right[0], dtype=rho.dtype)
        yr = np.full((b, 1), right[1], dtype=rho.dtype)
    else:
        raise ValueError("Unknown boundary condition: {}".format(
            boundary_conditions))
    rho_tp1 = np.concatenate((rl, rho_in, rr), axis=1)
    y_tp1 = np.concatenate((yl, y_in, yr), axis=1)

    # u(): in-place zero patch, then y/rho + V(rho) (arz.py:258-277).
    rho_tp1 = np.where(rho_tp1 == 0, ZERO_DENSITY_PATCH, rho_tp1).astype(
        rho_tp1.dtype)
    with np.errstate(divide="ignore", invalid="ignore"):
        v_tp1 = (y_tp1 / rho_tp1) + v_eq(rho_tp1, v_max, rho_max, lam)
    out = np.concatenate((rho_tp1, v_tp1), axis=1)
    return out[0] if squeeze else out


# --------------------------------------------------------------------------- #
# Non-local look-ahead model -- no reference code; parity unpinned            #
# --------------------------------------------------------------------------- #

def _two_sum(x, y):
    s = x + y
    bb = s - x
    return s, (x - (s - bb)) + (y - bb)


def _veltkamp(a):
    c = 134217729.0 * a          # 2**27 + 1
    hi = c - (c - a)
    return hi, a - hi


def fma(a, b, c):
    """Correctly rounded ``a * b + c`` (one rounding), element-wise, float64.

    NumPy has no fused multiply-add; this is the classic error-free route:
    Dekker's exact product ``p + e = a * b`` (Veltkamp splitting), ``s + t = p +
    c`` exactly (Knuth's TwoSum), the small terms ``t + e`` rounded TO ODD, and
    one final round-to-nearest addition (Boldo & Melquiond, "Emulation of FMA and
    correctly rounded sums", IEEE TC 2008).  Valid away from overflow and
    underflow, which is where the traffic models live; checked against the C
    library's ``fma`` in tests/test_nonlocal_unpinned.py.  float32 inputs are
    evaluated in float64 and rounded once more (the non-local float32 tests are
    tolerance tests)."""
    out_dtype = np.result_type(a, b, c)
    a, b, c = np.broadcast_arrays(np.asarray(a, np.float64), np.asarray(b, np.float64),
                                  np.asarray(c, np.float64))
    p = a * b
    ah, al = _veltkamp(a)
    bh, bl = _veltkamp(b)
    e = ((ah * bh - p) + ah * bl + al * bh) + al * bl
    s, t = _two_sum(p, c)
    u, r = _two_sum(t, e)
    even = (np.ascontiguousarray(u).view(np.int64) & 1) == 0
    v = np.where((r != 0) & even, np.nextafter(u, np.where(r > 0, np.inf, -np.inf)), u)
    res = s + v
    return res.astype(out_dtype) if out_dtype == np.float32 else res


def nonlocal_weights(eta, dx, kernel="linear", dtype=np.float64):
    """Cell weights gamma_k = int_{k dx}^{(k+1) dx} w_eta(y) dy, k < n_eta.

    Friedrich, Kolb, Goettlich (2018), the paper cited at lwr.py:209-211:
    ``w_eta`` is non-increasing on ``[0, eta]`` with unit mass.  ``linear``:
    w(y) = 2 (eta - y) / eta^2;  ``constant``: w(y) = 1 / eta.  ``eta`` must
    be a whole number of cells.
    """
    n_eta = int(round(eta / dx))
    if n_eta < 1 or abs(n_eta * dx - eta) > 1e-9 * max(1.0, abs(eta)):
        raise ValueError("eta must be a positive multiple of dx")
    k = np.arange(n_eta, dtype=np.float64)
    a, b = k * dx, (k + 1) * dx
    if kernel == "linear":
        g = (2.0 * eta * (b - a) - (b * b - a * a)) / (eta * eta)
    elif kernel == "constant":
        g = (b - a) / eta
    else:
        raise ValueError("Unknown look-ahead kernel: {}".format(kernel))
    return g.astype(dtype)


def _nonlocal_ghosts(rho, n_eta, boundary_conditions):
    """Pad 1 upstream ghost and n_eta downstream ghosts."""
    b = rho.shape[0]
    bc = boundary_conditions
    if bc == "loop":  # a true periodic ring for this model
        left = rho[:, -1:]
        idx = np.arange(n_eta) % rho.shape[1]
        right = rho[:, idx]
    elif bc == "extend_both":
        left = rho[:, :1]
        right = np.repeat(rho[:, -1:], n_eta, axis=1)
    elif isinstance(bc, dict) and list(bc.keys())[0] == "constant_both":
        lo, hi = bc["constant_both"]
        left = np.full((b, 1), lo, dtype=rho.dtype)
        right = np.full((b, n_eta), hi, dtype=rho.dtype)
    elif isinstance(bc, dict) and list(bc.keys())[0] == "inflow_outflow":
        # fixed inflow density upstream, free outflow downstream (config 3)
        lo = bc["inflow_outflow"]
        left = np.full((b, 1), lo, dtype=rho.dtype)
        right = np.repeat(rho[:, -1:], n_eta, axis=1)
    else:
        raise ValueError("Unknown boundary conditions: {}".format(bc))
    return np.concatenate((left, rho, right), axis=1)


def nonlocal_step(obs, gamma, dx=50, dt=0.5, rho_max=1, v_max=30, lam=1,
                  boundary_conditions="loop"):
    """Godunov-type step of the non-local (mean downstream velocity) LWR model.

    Scheme (2.4)-(2.5) of Friedrich, Kolb, Goettlich, NHM 13 (2018) with
    g(rho) = rho:

        W_j     = sum_{k=0}^{n_eta-1} gamma_k V(rho_{j+k+1})   (k ascending,
                  each term joined by a fused multiply-add: w = fma(gamma_k, V, w)
                  -- one rounding per term; no reference code fixes the order)
        F_{j+1/2} = rho_j W_j
        rho'_j  = rho_j - (dt/dx) (F_{j+1/2} - F_{j-1/2})

    The observation keeps the reference layout; like ``LWRModel`` the input
    velocity half is ignored and the output velocity half is V(rho').
    """
    rho, _, squeeze = _split(obs)
    gamma = np.asarray(gamma, dtype=rho.dtype)
    n_eta = gamma.shape[0]
    n = rho.shape[1]
    step = dt / dx
    ext = _nonlocal_ghosts(rho, n_eta, boundary_conditions)  # (B, 1+N+n_eta)
    ve = v_eq(ext, v_max, rho_max, lam)
    # W for cells -1 .. N-1  (ext index j+1 <-> cell j)
    w = np.zeros((rho.shape[0], n + 1), dtype=rho.dtype)
    for k in range(n_eta):
        w = fma(gamma[k], ve[:, k + 1:k + 1 + n + 1], w)
    flux = ext[:, :n + 1] * w            # F_{j+1/2} for j = -1 .. N-1
    rho_tp1 = rho - step * (flux[:, 1:] - flux[:, :-1])
    v_tp1 = v_eq(rho_tp1, v_max, rho_max, lam)
    out = np.concatenate((rho_tp1, v_tp1), axis=1)
    return out[0] if squeeze else out


def nonlocal_lxf_step(obs, gamma, dx=50, dt=0.5, rho_max=1, v_max=30, lam=1,
                      boundary_conditions="loop", alpha=None):
    """Lax-Friedrichs-type step of the non-local LWR model whose speed depends
    on a downstream average of the DENSITY (Blandin & Goatin, Numer. Math. 132
    (2016), scheme (3.2)-(3.4); the second scheme BASELINE.json's north_star
    names).  PARITY UNPINNED: the reference holds no code for it.

        R_j       = sum_{k=0}^{n_eta-1} gamma_k rho_{j+k}        (k ascending, fused
                    multiply-adds as in nonlocal_step)
        g_j       = rho_j V(R_j)
        F_{j+1/2} = 0.5 (g_j + g_{j+1}) + (0.5 alpha) (rho_j - rho_{j+1})
        rho'_j    = rho_j - (dt/dx) (F_{j+1/2} - F_{j-1/2})

    ``alpha`` is the numerical viscosity, >= max |V| (default: v_max; the
    scheme is monotone for dt/dx <= 1 / (alpha + gamma_0 |V'| rho_max)-ish
    step sizes, see the paper's CFL condition (3.5)).  Ghost cells as in
    ``nonlocal_step`` (one upstream, n_eta downstream).  The weights are the
    cell integrals ``nonlocal_weights`` returns (the paper uses left-endpoint
    values w(k dx) dx; both are consistent quadratures of the same kernel).
    """
    rho, _, squeeze = _split(obs)
    gamma = np.asarray(gamma, dtype=rho.dtype)
    n_eta = gamma.shape[0]
    n = rho.shape[1]
    step = dt / dx
    ext = _nonlocal_ghosts(rho, n_eta, boundary_conditions)   # cells -1 .. N+n_eta-1
    # R for cells -1 .. N  (ext index c+1 <-> cell c)
    r = np.zeros((rho.shape[0], n + 2), dtype=rho.dtype)
    for k in range(n_eta):
        r = fma(gamma[k], ext[:, k:k + n + 2], r)
    cells = ext[:, :n + 2]
    g = cells * v_eq(r, v_max, rho_max, lam)
    half_alpha = 0.5 * _col(v_max if alpha is None else alpha, rho.dtype)
    flux = 0.5 * (g[:, :-1] + g[:, 1:]) + half_alpha * (cells[:, :-1] - cells[:, 1:])
    rho_tp1 = rho - step * (flux[:, 1:] - flux[:, :-1])       # faces -1/2 .. N-1/2
    v_tp1 = v_eq(rho_tp1, v_max, rho_max, lam)
    out = np.concatenate((rho_tp1, v_tp1), axis=1)
    return out[0] if squeeze else out


# --------------------------------------------------------------------------- #
# drivers                                                                     #
# --------------------------------------------------------------------------- #

def rollout(step_fn, obs, n_steps, **params):
    """Chain ``n_steps`` calls of ``step_fn`` (evaluate_model.py:398-403)."""
    cur = np.array(obs, copy=True)
    for _ in range(n_steps):
        cur = step_fn(cur, **params)
    return cur


def density_mse(next_states, predicted):
    """Loss of ``compute_loss``: MSE over the density half only
    (lwr.py:338-343, arz.py:515-520; sklearn ``mean_squared_error``)."""
    n = int(round(np.shape(next_states)[1] / 2))
    a = np.asarray(next_states)[:, :n]
    b = np.asarray(predicted)[:, :n]
    return float(np.mean(np.mean((a - b) ** 2, axis=0)))


def mean_speed_reward(obs):
    """Density-weighted mean speed sum(rho v) / sum(rho) per row -- the
    macroscopic analogue of ``VSLEnv.compute_reward`` (envs/vsl.py:115-118)."""
    rho, v, squeeze = _split(obs)
    r = (rho * v).sum(axis=1) / rho.sum(axis=1)
    return r[0] if squeeze else r


def aggregate_micro(time, position, speed, dx=50.0, dt=0.5, length=1500.0,
                    total_time=1800.0, empty_speed=30.0):
    """Micro -> macro aggregation exactly as the reference's notebook loop
    (experiments/plotting_results.ipynb cell 8), record by record."""
    import math
    n_rows, n_sec = int(round(total_time / dt)) + 1, int(round(length / dx))
    speeds = np.zeros((n_rows, n_sec))
    num = np.zeros((n_rows, n_sec))
    for t, x, v in zip(time, position, speed):
        ix, iy = math.floor(float(t) / dt), math.floor(float(x) / dx)
        try:
            speeds[ix, iy] += float(v)
            num[ix, iy] += 1
        except IndexError:
            pass
    dens = num / dx
    speeds[num == 0] = empty_speed
    safe = num.copy()
    safe[safe == 0] = 1
    return speeds / safe, dens
