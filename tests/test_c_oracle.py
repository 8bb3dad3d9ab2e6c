"""The C restatement (CPU baseline) against the NumPy oracle and the golden
vectors: bit-equal for lam in {1, 2, 0.5}."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import c_oracle, np_oracle as o

CASES = load_golden()


@pytest.mark.parametrize("case", CASES, ids=[c.key for c in CASES])
def test_c_oracle_matches_numpy_oracle_on_golden_inputs(case):
    kw = case.step_kwargs()
    fn = o.lwr_step if case.model == "lwr" else o.arz_step
    ref = fn(case.obs, **kw)
    got = c_oracle.step(case.model, case.obs, **kw)
    assert np.array_equal(got, ref)
    if case.model == "lwr":
        assert np.array_equal(got, case.out1)      # the reference itself
    cur_c, cur_n = case.obs, case.obs
    for _ in range(20):
        cur_c = c_oracle.step(case.model, cur_c, **kw)
        cur_n = fn(cur_n, **kw)
    assert np.array_equal(cur_c, cur_n)


@pytest.mark.parametrize("model", ["lwr", "arz"])
@pytest.mark.parametrize("lam", [1, 2, 0.5])
def test_c_oracle_batch_and_threads(model, lam):
    rng = np.random.default_rng(5)
    b, n = 64, 257
    rho = rng.uniform(0.05, 0.9, (b, n))
    v = 30 * (1 - rho) ** lam + rng.normal(0, 1, (b, n))
    obs = np.concatenate((rho, v), axis=1)
    fn = o.lwr_step if model == "lwr" else o.arz_step
    ref = fn(obs, lam=lam)
    for threads in (1, 4):
        got = c_oracle.step(model, obs, lam=lam, n_threads=threads)
        assert np.array_equal(got, ref, equal_nan=True)
