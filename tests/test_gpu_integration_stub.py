"""Runs the ctypes stub printed in INTEGRATION.md section 2 -- the binding a
reference maintainer would paste into mbrl_traffic/models/arz.py -- exactly as
printed, against the oracle."""
import os
import re

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub_source():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = [b for b in blocks if "mt_step_host" in b and "class MtParams" in b]
    assert len(stub) == 1, "INTEGRATION.md section 2 must hold exactly one stub"
    return stub[0]


def test_the_integration_stub_runs_as_printed(monkeypatch):
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (no CPU fallback exists)")
    from oracle import np_oracle as o
    monkeypatch.setenv("MT_LIB", os.path.join(ROOT, "mbrl_traffic_b200", "libmtraffic.so"))

    class Model(object):          # stands in for mbrl_traffic.models.base.Model
        pass

    ns = {"Model": Model}
    exec(compile(_stub_source(), "INTEGRATION.md#stub", "exec"), ns)
    m = ns["ARZModel"]()
    m.dx, m.dt, m.rho_max, m.v_max, m.tau, m.lam = 50.0, 0.5, 1.0, 30.0, 9.0, 1.0
    m.rho_crit = 0.5
    rng = np.random.default_rng(5)
    for n in (30, 1000):
        m._mt = None
        rho = rng.uniform(0.05, 0.9, n)
        obs = np.concatenate((rho, 30 * (1 - rho) + rng.normal(0, 1, n)))
        keep = obs.copy()
        out = m.get_next_obs(obs, None)
        assert np.array_equal(obs, keep)            # the input is not modified
        assert np.array_equal(out, o.arz_step(obs[None])[0])
        out2 = m.get_next_obs(out, None)            # the handle is reused
        assert np.array_equal(out2, o.arz_step(out[None])[0])
