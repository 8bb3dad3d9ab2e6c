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
