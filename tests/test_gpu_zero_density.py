"""MT_FLAG_FSOLVE_ABORT on the GPU: roads with an exact-zero density follow the
reference, whose fsolve call gives up there and leaves the relative flow frozen
for the whole road (arz.py:411, 460-463).  Checked against outputs of the
unmodified reference (tests/golden/reference_zero_density.npz) and, for
batches, against the oracle's restatement of the rule."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import np_oracle as o  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (no CPU fallback exists)")


def _model(p, **extra):
    from mbrl_traffic_b200 import ARZModel
    return ARZModel(None, None, None, None, 0, dx=p["dx"], dt=p["dt"], rho_max=p["rho_max"],
                    rho_max_max=1, v_max=p["v_max"], v_max_max=30, tau=p["tau"], tau_max=25,
                    lam=p["lam"], boundary_conditions=p["boundary_conditions"],
                    optimizer_cls=None, **extra)


def test_zero_density_roads_equal_the_reference_outputs():
    _need_gpu()
    z = np.load(os.path.join(HERE, "golden", "reference_zero_density.npz"))
    keys = sorted(set(k.split("/")[0] for k in z.files))
    assert len(keys) == 88
    for key in keys:
        meta = json.loads(str(z[key + "/meta"]))
        p, n = meta["params"], meta["n"]
        if isinstance(p["boundary_conditions"], dict):     # JSON turned the tuples into lists
            p["boundary_conditions"] = {k: tuple(tuple(x) for x in v)
                                        for k, v in p["boundary_conditions"].items()}
        obs, ref = z[key + "/obs"], z[key + "/out1"]
        m = _model(p, fsolve_abort=True)
        got = m.get_next_obs(torch.as_tensor(obs, device="cuda:0"), None).cpu().numpy()
        assert np.array_equal(got[:n], ref[:n]), key
        np.testing.assert_allclose(got[n:], ref[n:], rtol=1e-12, atol=0, err_msg=key)
        host = m.get_next_obs(obs, None)                   # host-buffer path
        assert np.array_equal(host, got), key
        m.close()


@pytest.mark.parametrize("n,bc", [(1000, "loop"), (1000, "extend_both"), (250, "loop"),
                                  (33, "extend_both"), (40000, "extend_both")])
def test_mixed_batch_bit_equal_to_the_oracle_rule(n, bc):
    """Some roads hold zeros, most do not; rewards and the bfloat16 copy follow."""
    _need_gpu()
    from mbrl_traffic_b200.engine import Engine
    b = 64 if n <= 1000 else 3
    rng = np.random.default_rng(n)
    rho = rng.uniform(0.05, 0.9, (b, n))
    v = 30 * (1 - rho) + rng.normal(0, 1, (b, n))
    for road in range(0, b, 3):
        rho[road, rng.integers(0, n, size=1 + road % 3)] = 0.0
    obs = np.concatenate((rho, v), axis=1)
    P = dict(dx=50, dt=0.5, rho_max=1, v_max=30, tau=9, lam=1, boundary_conditions=bc)
    ref = o.arz_step(obs, fsolve_abort=True, **P)
    assert np.isfinite(ref).all()
    assert not np.isfinite(o.arz_step(obs, **P)).all()      # the plain rule leaves NaNs
    eng = Engine("arz", "float64", b, n)
    eng.set_params(fsolve_abort=True, **P)
    x = torch.as_tensor(obs, device="cuda:0")
    y = torch.empty_like(x)
    rew = torch.empty(b, dtype=torch.float64, device="cuda:0")
    o16 = torch.empty(b, 2 * n, dtype=torch.bfloat16, device="cuda:0")
    eng.step(x, y, reward=rew, obs16=o16)
    assert np.array_equal(y.cpu().numpy(), ref)
    np.testing.assert_allclose(rew.cpu().numpy(), o.mean_speed_reward(ref), rtol=1e-12)
    assert torch.equal(o16, y.to(torch.bfloat16))
    # several steps: mt_rollout falls back to single steps under the flag
    ref3 = obs
    for _ in range(3):
        ref3 = o.arz_step(ref3, fsolve_abort=True, **P)
    out3 = eng.rollout(x.clone(), y, 3)
    assert np.array_equal(out3.cpu().numpy(), ref3)
    eng.close()


def test_flag_off_keeps_the_closed_form_and_f32_follows_the_rule():
    _need_gpu()
    from mbrl_traffic_b200.engine import Engine
    b, n = 8, 1000
    rng = np.random.default_rng(2)
    rho = rng.uniform(0.05, 0.9, (b, n))
    v = 30 * (1 - rho) + rng.normal(0, 1, (b, n))
    rho[1, 500] = 0.0
    obs = np.concatenate((rho, v), axis=1)
    P = dict(dx=50, dt=0.5, rho_max=1, v_max=30, tau=9, lam=1, boundary_conditions="loop")
    eng = Engine("arz", "float64", b, n)
    eng.set_params(**P)
    x = torch.as_tensor(obs, device="cuda:0")
    y = torch.empty_like(x)
    eng.step(x, y)
    assert np.array_equal(y.cpu().numpy(), o.arz_step(obs, **P), equal_nan=True)
    eng.close()
    e32 = Engine("arz", "float32", b, n)
    e32.set_params(fsolve_abort=True, **P)
    x32 = torch.as_tensor(obs.astype(np.float32), device="cuda:0")
    y32 = torch.empty_like(x32)
    e32.step(x32, y32)
    ref = o.arz_step(obs.astype(np.float32).astype(np.float64), fsolve_abort=True, **P)
    err = np.abs(y32.cpu().numpy() - ref) / np.maximum(np.abs(ref), 1.0)
    assert np.isfinite(y32.cpu().numpy()).all() and err.max() <= 2e-6
    e32.close()
