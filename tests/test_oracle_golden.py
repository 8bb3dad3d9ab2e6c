"""The oracle against the golden vectors produced by the reference itself.

The reference's own tests leave the numeric step untested
(tests/fast_tests/test_models.py:74-98, 213-237), so the vectors in
tests/golden were produced by executing the unmodified reference
(oracle/make_golden.py).  Bars: LWR bit-equal; ARZ density half bit-equal after
one step, velocity half within 1e-12 relative (fsolve noise).
"""
import os

import numpy as np
import pytest

from conftest import load_golden
from oracle import np_oracle as o
from oracle import ref_shim

CASES = load_golden()
IDS = [c.key for c in CASES]


def _step(case, obs):
    fn = o.lwr_step if case.model == "lwr" else o.arz_step
    return fn(obs, **case.step_kwargs())


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_one_step_matches_reference(case):
    out = _step(case, case.obs)
    assert out.shape == case.out1.shape
    n = out.shape[-1] // 2
    if case.model == "lwr":
        assert np.array_equal(out, case.out1)
    else:
        assert np.array_equal(out[..., :n], case.out1[..., :n])
        np.testing.assert_allclose(out[..., n:], case.out1[..., n:],
                                   rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_chained_steps_match_reference(case):
    cur = case.obs
    for _ in range(case.n_chain):
        cur = _step(case, cur)
    if case.model == "lwr":
        assert np.array_equal(cur, case.outn)
    else:
        # fsolve's few-ulp noise feeds back through the chain; the schemes
        # are dissipative so it stays tiny (worst probe 1.2e-10 relative).
        np.testing.assert_allclose(cur, case.outn, rtol=1e-8, atol=1e-10)


def test_batched_equals_row_by_row():
    rng = np.random.default_rng(7)
    b, n = 5, 24
    rho = rng.uniform(0.05, 0.9, (b, n))
    v = 30 * (1 - rho) + rng.normal(0, 1, (b, n))
    obs = np.concatenate((rho, v), axis=1)
    for fn, kw in ((o.lwr_step, {}), (o.arz_step, {"tau": 9})):
        full = fn(obs, **kw)
        for i in range(b):
            assert np.array_equal(full[i], fn(obs[i], **kw))


def test_per_env_parameters_equal_separate_models():
    rng = np.random.default_rng(8)
    b, n = 4, 16
    rho_max = np.array([1.0, 0.2, 0.5, 1.0])
    v_max = np.array([30.0, 25.0, 20.0, 10.0])
    lam = np.array([1.0, 2.0, 1.0, 0.5])
    tau = np.array([9.0, 5.0, 2.0, 20.0])
    rho = rho_max[:, None] * rng.uniform(0.05, 0.9, (b, n))
    v = rng.uniform(1, 20, (b, n))
    obs = np.concatenate((rho, v), axis=1)
    full = o.arz_step(obs, rho_max=rho_max, v_max=v_max, lam=lam, tau=tau)
    full_l = o.lwr_step(obs, rho_max=rho_max, v_max=v_max, lam=lam)
    for i in range(b):
        one = o.arz_step(obs[i], rho_max=float(rho_max[i]),
                         v_max=float(v_max[i]), lam=float(lam[i]),
                         tau=float(tau[i]))
        assert np.array_equal(full[i], one)
        one = o.lwr_step(obs[i], rho_max=float(rho_max[i]),
                         v_max=float(v_max[i]), lam=float(lam[i]))
        assert np.array_equal(full_l[i], one)


def test_unknown_boundary_condition_raises():
    obs = np.full(16, 0.5)
    with pytest.raises(ValueError, match="Unknown boundary conditions"):
        o.lwr_step(obs, boundary_conditions="ring")
    with pytest.raises(ValueError, match="Unknown boundary condition"):
        o.arz_step(obs, boundary_conditions="ring")


def test_uniform_state_is_a_fixed_point():
    n = 32
    rho = np.full(n, 0.3)
    v = 30 * (1 - rho)
    obs = np.concatenate((rho, v))
    for bc in ("loop", "extend_both"):
        np.testing.assert_allclose(
            o.lwr_step(obs, boundary_conditions=bc), obs, rtol=0, atol=1e-14)
        np.testing.assert_allclose(
            o.arz_step(obs, boundary_conditions=bc), obs, rtol=0, atol=1e-13)
    g = o.nonlocal_weights(200.0, 50.0)
    np.testing.assert_allclose(
        o.nonlocal_step(obs, g, boundary_conditions="loop"), obs,
        rtol=0, atol=1e-13)


def test_nonlocal_weights_and_conservation():
    g = o.nonlocal_weights(800.0, 50.0, "linear")
    assert g.shape == (16,)
    assert abs(g.sum() - 1.0) < 1e-14
    assert np.all(np.diff(g) < 0)
    gc = o.nonlocal_weights(800.0, 50.0, "constant")
    np.testing.assert_allclose(gc, 1.0 / 16)
    rng = np.random.default_rng(3)
    rho = rng.uniform(0.05, 0.9, (3, 64))
    obs = np.concatenate((rho, np.zeros_like(rho)), axis=1)
    cur = obs
    for _ in range(200):
        cur = o.nonlocal_step(cur, g, boundary_conditions="loop")
        r = cur[:, :64]
        assert r.min() >= 0 and r.max() <= 1.0   # max principle under CFL
    np.testing.assert_allclose(cur[:, :64].sum(axis=1), rho.sum(axis=1),
                               rtol=1e-12)        # ring conserves mass


def test_nonlocal_reduces_to_upwind_when_one_cell_lookahead():
    rng = np.random.default_rng(4)
    n = 40
    rho = rng.uniform(0.05, 0.9, n)
    obs = np.concatenate((rho, np.zeros(n)))
    out = o.nonlocal_step(obs, np.array([1.0]), boundary_conditions="loop")
    ve = 30 * (1 - rho)
    flux = rho * np.roll(ve, -1)          # rho_j V(rho_{j+1})
    exp = rho - 0.01 * (flux - np.roll(flux, 1))
    np.testing.assert_allclose(out[:n], exp, rtol=0, atol=1e-14)


@pytest.mark.skipif(not ref_shim.available(),
                    reason="reference checkout only exists in the build box")
def test_oracle_against_live_reference():
    rng = np.random.default_rng(11)
    n = 30
    rho = rng.uniform(0.05, 0.9, n)
    v = 30 * (1 - rho) + rng.normal(0, 1, n)
    obs = np.concatenate((rho, v))
    m = ref_shim.reference_lwr(dx=50, dt=0.5, rho_max=1, rho_max_max=1,
                               v_max=30, v_max_max=30, lam=1,
                               boundary_conditions="loop")
    assert np.array_equal(m.get_next_obs(obs.copy(), None), o.lwr_step(obs))
    m = ref_shim.reference_arz(dx=50, dt=0.5, rho_max=1, rho_max_max=1,
                               v_max=30, v_max_max=30, tau=9, tau_max=25,
                               lam=1, boundary_conditions="loop")
    ref = m.get_next_obs(obs.copy(), None)
    out = o.arz_step(obs)
    assert np.array_equal(ref[:n], out[:n])
    np.testing.assert_allclose(ref[n:], out[n:], rtol=1e-12)


# --------------------------------------------------------------------------- #
# exact-zero densities: the reference's fsolve gives up (arz.py:411, 460-463)  #
# --------------------------------------------------------------------------- #
def _zero_cases():
    import json
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                        "reference_zero_density.npz")
    z = np.load(path)
    for key in sorted(set(k.split("/")[0] for k in z.files)):
        yield key, z[key + "/obs"], z[key + "/out1"], json.loads(str(z[key + "/meta"]))


def test_oracle_mimics_the_reference_on_roads_with_an_exact_zero_density():
    """88 outputs of the unmodified reference (oracle/make_golden_zero.py): a
    road with rho == 0 somewhere keeps its relative flow for that step."""
    n_abort = 0
    for key, obs, ref, meta in _zero_cases():
        p, n = meta["params"], meta["n"]
        kw = dict(dx=p["dx"], dt=p["dt"], rho_max=p["rho_max"], v_max=p["v_max"], tau=p["tau"],
                  lam=p["lam"], boundary_conditions=p["boundary_conditions"])
        got = o.arz_step(obs, fsolve_abort=True, **kw)
        assert np.isfinite(ref).all(), key           # the reference never returns NaN here
        assert np.array_equal(got[:n], ref[:n]), key
        np.testing.assert_allclose(got[n:], ref[n:], rtol=1e-12, atol=0, err_msg=key)
        plain = o.arz_step(obs, **kw)
        if not np.array_equal(plain, got, equal_nan=True):
            n_abort += 1
    assert n_abort >= 40          # most cases do exercise the abort rule
