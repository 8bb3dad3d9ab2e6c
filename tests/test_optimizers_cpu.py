"""Population optimizers on an analytic objective (no GPU): the contract of
the reference's Optimizer base class (utils/optimizers/base.py:4-63) and
convergence inside the parameter box."""
import numpy as np
import pytest

from mbrl_traffic_b200.optimizers import (CrossEntropyMethod, GeneticAlgorithm,
                                          Optimizer)

TARGET = np.array([0.3, 21.0, 1.4])


def fitness(x):
    return float(np.sum(((np.asarray(x) - TARGET) / np.array([1, 30, 4])) ** 2))


def batch_fitness(pop):
    return np.sum(((pop - TARGET) / np.array([1, 30, 4])) ** 2, axis=1)


def test_base_contract():
    opt = Optimizer([0, 0], [1, 2], fitness, verbose=0)
    assert opt.param_low == [0, 0] and opt.param_high == [1, 2]
    assert opt.fitness_fn is fitness and opt.verbose == 0
    with pytest.raises(NotImplementedError):
        opt.solve()


@pytest.mark.parametrize("batched", [False, True])
def test_cem_converges(batched):
    opt = CrossEntropyMethod([0, 0, 0], [1, 30, 4], fitness, verbose=0,
                             population_size=64, seed=1,
                             batch_fitness_fn=batch_fitness if batched else None)
    x, f = opt.solve(num_steps=60)
    assert f < 1e-8
    np.testing.assert_allclose(x, TARGET, atol=1e-3)
    assert np.all(x >= 0) and np.all(x <= [1, 30, 4])


def test_ga_improves_and_respects_bounds():
    opt = GeneticAlgorithm([0, 0, 0], [1, 30, 4], fitness, verbose=0,
                           population_size=100, mutation_prob=0.2,
                           mutation_std=0.05, seed=2,
                           batch_fitness_fn=batch_fitness)
    x, f = opt.solve(num_steps=80)
    assert f < 1e-3
    assert np.all(x >= 0) and np.all(x <= [1, 30, 4])
    with pytest.raises(ValueError):
        GeneticAlgorithm([0], [1], fitness, selection_method="roulette")


def test_termination_fn_stops_early():
    calls = []
    opt = CrossEntropyMethod([0, 0, 0], [1, 30, 4], fitness, verbose=0, seed=3,
                             batch_fitness_fn=lambda p: (calls.append(1),
                                                         batch_fitness(p))[1])
    opt.solve(num_steps=500, termination_fn=lambda f: f < 1e-2)
    assert len(calls) < 100


# --------------------------------------------------------------------------- #
# batched Nelder-Mead (reference: utils/optimizers/nelder_mead.py:81-95)       #
# --------------------------------------------------------------------------- #

def _rosen(x):
    x = np.asarray(x, dtype=np.float64)
    return float(100.0 * (x[1] - x[0] ** 2) ** 2 + (1 - x[0]) ** 2 + (x[2] - 0.5) ** 2)


def _rosen_batch(pop):
    return np.array([_rosen(x) for x in pop])


@pytest.mark.parametrize("batched", [False, True])
@pytest.mark.parametrize("adaptive", [False, True])
def test_nelder_mead_visits_the_same_simplices_as_scipy(batched, adaptive):
    """One simplex, speculative batched scoring of the four trial points: the
    accepted points are the ones scipy's sequential rules accept, so after the
    same number of iterations the best vertex is the same."""
    from scipy import optimize
    from mbrl_traffic_b200.optimizers import NelderMead
    low, high = [-2.0, -1.0, 0.0], [2.0, 3.0, 1.0]
    x0 = [-1.2, 1.0, 0.9]
    for iters in (5, 40, 200):
        ref = optimize.minimize(_rosen, x0, method="Nelder-Mead",
                                bounds=list(zip(low, high)),
                                options=dict(maxiter=iters, xatol=1e-10, fatol=1e-12,
                                             adaptive=adaptive, disp=False))
        opt = NelderMead(low, high, _rosen, verbose=0, x0=x0, disp=False,
                         xatol=1e-10, fatol=1e-12, adaptive=adaptive,
                         batch_fitness_fn=_rosen_batch if batched else None)
        x, f = opt.solve(num_steps=iters)
        np.testing.assert_allclose(x, ref.x, rtol=0, atol=1e-12)
        assert abs(f - ref.fun) <= 1e-12 * max(1.0, abs(ref.fun))


def test_nelder_mead_reference_constructor_and_multi_start():
    from mbrl_traffic_b200.optimizers import NelderMead
    opt = NelderMead([0, 0, 0], [1, 30, 4], fitness, verbose=0)
    # the reference's defaults (nelder_mead.py:22-79)
    np.testing.assert_allclose(opt.x0, [0.5, 15.0, 2.0])
    assert opt.options == {"disp": True, "initial_simplex": None, "xatol": 1e-8,
                           "fatol": 1000, "adaptive": False}
    assert opt.bnds == ([0, 0, 0], [1, 30, 4])
    calls = []

    def counted(pop):
        calls.append(len(pop))
        return batch_fitness(pop)

    opt = NelderMead([0, 0, 0], [1, 30, 4], fitness, verbose=0, fatol=1e-14,
                     n_starts=8, batch_fitness_fn=counted, seed=1)
    x, f = opt.solve(num_steps=300)
    np.testing.assert_allclose(x, TARGET, atol=1e-4)
    assert f < 1e-9
    # every iteration scored the four trial points of all 8 simplices together
    assert calls[0] == 8 * 4 and 32 in calls[1:]
    hit = []
    opt = NelderMead([0, 0, 0], [1, 30, 4], fitness, verbose=0, fatol=1e-14,
                     batch_fitness_fn=batch_fitness)
    opt.solve(num_steps=500, termination_fn=lambda best: hit.append(best) or best < 1e-3)
    assert hit[-1] < 1e-3 and len(hit) < 500
