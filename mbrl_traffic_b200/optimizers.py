"""Population-based optimizers for the model parameters.

Same ``Optimizer`` contract as the reference
(``mbrl_traffic/utils/optimizers/base.py:4-63``: ``__init__(param_low,
param_high, fitness_fn, verbose)``, ``solve(num_steps, termination_fn) ->
(params, fitness)``).  The reference's ``CrossEntropyMethod.solve`` is ``pass``
(``cross_entropy.py:34-36``) and its ``GeneticAlgorithm`` is half-written
(``genetic_algorithm.py:120-156``); these are working versions whose whole
population is scored by ONE batched model step: every candidate parameter
vector becomes the per-road parameters of its own block of roads
(``FiniteVolumeModel.population_loss``), instead of the reference's
candidates x rows x iterations Python loops (``lwr.py:140-177, 329-343``).
"""
import numpy as np


class Optimizer(object):
    """Base optimizer object (reference ``optimizers/base.py:4-63``)."""

    #: the models hand ``batch_fitness_fn`` to optimizers that set this
    accepts_batch_fitness = False

    def __init__(self, param_low, param_high, fitness_fn, verbose=2):
        self.param_low = param_low
        self.param_high = param_high
        self.fitness_fn = fitness_fn
        self.verbose = verbose

    def solve(self, num_steps=1000, termination_fn=None):
        raise NotImplementedError


class _PopulationOptimizer(Optimizer):
    accepts_batch_fitness = True

    def __init__(self, param_low, param_high, fitness_fn, verbose=2,
                 population_size=64, batch_fitness_fn=None, seed=None):
        super(_PopulationOptimizer, self).__init__(
            param_low=param_low, param_high=param_high,
            fitness_fn=fitness_fn, verbose=verbose)
        self.population_size = population_size
        self.batch_fitness_fn = batch_fitness_fn
        self.rng = np.random.default_rng(seed)
        self.low = np.asarray(param_low, dtype=np.float64)
        self.high = np.asarray(param_high, dtype=np.float64)

    def _score(self, population):
        """Fitness (loss; lower is better) of every row of ``population``."""
        if self.batch_fitness_fn is not None:
            return np.asarray(self.batch_fitness_fn(population), dtype=np.float64)
        return np.array([self.fitness_fn(list(x)) for x in population],
                        dtype=np.float64)


class CrossEntropyMethod(_PopulationOptimizer):
    """Cross-entropy method: sample a Gaussian population inside the box,
    keep the elite fraction, refit mean and spread, repeat."""

    def __init__(self, param_low, param_high, fitness_fn, verbose=2,
                 population_size=64, elite_frac=0.125, init_std_frac=0.3,
                 min_std_frac=1e-6, smoothing=0.0, batch_fitness_fn=None,
                 seed=None):
        super(CrossEntropyMethod, self).__init__(
            param_low, param_high, fitness_fn, verbose, population_size,
            batch_fitness_fn, seed)
        self.n_elite = max(2, int(round(elite_frac * population_size)))
        self.init_std_frac = init_std_frac
        self.min_std_frac = min_std_frac
        self.smoothing = smoothing

    def solve(self, num_steps=1000, termination_fn=None):
        span = self.high - self.low
        mean = 0.5 * (self.low + self.high)
        std = self.init_std_frac * span
        best_x, best_f = mean.copy(), np.inf
        for it in range(num_steps):
            pop = mean + std * self.rng.standard_normal(
                (self.population_size, mean.size))
            pop = np.clip(pop, self.low, self.high)
            pop[0] = mean                     # keep the incumbent in the race
            f = self._score(pop)
            f = np.where(np.isfinite(f), f, np.inf)
            order = np.argsort(f)
            if f[order[0]] < best_f:
                best_f, best_x = float(f[order[0]]), pop[order[0]].copy()
            elite = pop[order[:self.n_elite]]
            new_mean, new_std = elite.mean(axis=0), elite.std(axis=0)
            mean = self.smoothing * mean + (1 - self.smoothing) * new_mean
            std = np.maximum(self.smoothing * std + (1 - self.smoothing) * new_std,
                             self.min_std_frac * span)
            if self.verbose >= 2 and it % 10 == 0:
                print("CEM it %d best %.6g mean %s" % (it, best_f, mean))
            if termination_fn is not None and termination_fn(best_f):
                break
            if np.all(std <= self.min_std_frac * span * 1.0001):
                break
        return best_x, best_f


class GeneticAlgorithm(_PopulationOptimizer):
    """Genetic algorithm with the reference's knobs
    (``genetic_algorithm.py:49-118``): fittest selection of half the
    population, random pairing, uniform crossover, Gaussian mutation."""

    def __init__(self, param_low, param_high, fitness_fn, verbose=2,
                 population_size=100, selection_method="fittest",
                 pairing_method="random", mating_method="uniform",
                 mutation_method="gauss", mutation_prob=0.01,
                 mutation_std=0.001, batch_fitness_fn=None, seed=None):
        super(GeneticAlgorithm, self).__init__(
            param_low, param_high, fitness_fn, verbose, population_size,
            batch_fitness_fn, seed)
        for name, val, ok in (("selection_method", selection_method, ("fittest",)),
                              ("pairing_method", pairing_method, ("random",)),
                              ("mating_method", mating_method, ("uniform",)),
                              ("mutation_method", mutation_method, ("gauss",))):
            if val not in ok:
                raise ValueError("Unknown %s: %s" % (name, val))
        self.selection_method = selection_method
        self.pairing_method = pairing_method
        self.mating_method = mating_method
        self.mutation_method = mutation_method
        self.mutation_prob = mutation_prob
        self.mutation_std = mutation_std

    def solve(self, num_steps=1000, termination_fn=None):
        p, d = self.population_size, self.low.size
        span = self.high - self.low
        pop = self.rng.uniform(self.low, self.high, size=(p, d))
        best_x, best_f = pop[0].copy(), np.inf
        for it in range(num_steps):
            f = self._score(pop)
            f = np.where(np.isfinite(f), f, np.inf)
            order = np.argsort(f)
            if f[order[0]] < best_f:
                best_f, best_x = float(f[order[0]]), pop[order[0]].copy()
            if termination_fn is not None and termination_fn(best_f):
                break
            parents = pop[order[:p // 2]]                     # selection: fittest
            n_child = p - parents.shape[0]
            ia = self.rng.integers(0, parents.shape[0], n_child)   # pairing: random
            ib = self.rng.integers(0, parents.shape[0], n_child)
            mask = self.rng.random((n_child, d)) < 0.5        # mating: uniform
            children = np.where(mask, parents[ia], parents[ib])
            mutate = self.rng.random((n_child, d)) < self.mutation_prob
            children = children + mutate * self.rng.standard_normal(
                (n_child, d)) * self.mutation_std * span      # mutation: gauss
            pop = np.clip(np.concatenate((parents, children)), self.low, self.high)
            if self.verbose >= 2 and it % 10 == 0:
                print("GA it %d best %.6g" % (it, best_f))
        return best_x, best_f


class NelderMead(_PopulationOptimizer):
    """Nelder-Mead simplex search with the reference's constructor
    (``optimizers/nelder_mead.py:22-79``: ``x0``, ``disp``,
    ``initial_simplex``, ``xatol``, ``fatol``, ``adaptive``) and the moves and
    bound handling of ``scipy.optimize.minimize(method="Nelder-Mead",
    bounds=...)``, which the reference calls (``nelder_mead.py:81-95``).

    Batched: the reflection, expansion and both contraction points of every
    simplex are scored TOGETHER by one ``batch_fitness_fn`` call (one model
    step over ``4 * n_starts`` candidate blocks) instead of up to two
    sequential fitness calls; the accepted point is the one the sequential
    rules pick, so a run with ``n_starts=1`` visits the same simplices as
    SciPy.  ``n_starts > 1`` runs that many simplices side by side (the first
    from ``x0``, the others from seeded random points of the box) and returns
    the best vertex found -- the same kernel launch serves all of them.
    """

    def __init__(self, param_low, param_high, fitness_fn, verbose=2, x0=None,
                 disp=True, initial_simplex=None, xatol=1e-8, fatol=1000,
                 adaptive=False, n_starts=1, batch_fitness_fn=None, seed=None):
        super(NelderMead, self).__init__(
            param_low, param_high, fitness_fn, verbose, 4 * n_starts,
            batch_fitness_fn, seed)
        self.x0 = np.asarray(x0, dtype=np.float64) if x0 is not None \
            else 0.5 * (self.low + self.high)
        self.bnds = (param_low, param_high)
        self.n_starts = int(n_starts)
        self.options = {"disp": disp, "initial_simplex": initial_simplex,
                        "xatol": xatol, "fatol": fatol, "adaptive": adaptive}

    def _initial_simplices(self):
        d = self.low.size
        starts = [np.clip(self.x0, self.low, self.high)]
        for _ in range(self.n_starts - 1):
            starts.append(self.rng.uniform(self.low, self.high))
        sims = np.empty((self.n_starts, d + 1, d))
        for s, x0 in enumerate(starts):
            if s == 0 and self.options["initial_simplex"] is not None:
                sims[0] = np.asarray(self.options["initial_simplex"], dtype=np.float64)
                continue
            sims[s, 0] = x0                          # scipy: 5 % steps, 0.00025 from zero
            for k in range(d):
                y = x0.copy()
                y[k] = 1.05 * y[k] if y[k] != 0 else 0.00025
                sims[s, k + 1] = y
        return np.clip(sims, self.low, self.high)

    def solve(self, num_steps=1000, termination_fn=None):
        d, s_n = self.low.size, self.n_starts
        if self.options["adaptive"]:
            rho, chi, psi, sigma = 1.0, 1 + 2.0 / d, 0.75 - 1.0 / (2 * d), 1 - 1.0 / d
        else:
            rho, chi, psi, sigma = 1.0, 2.0, 0.5, 0.5
        xatol, fatol = self.options["xatol"], self.options["fatol"]
        sim = self._initial_simplices()                         # (S, d+1, d)
        fsim = self._score(sim.reshape(-1, d)).reshape(s_n, d + 1)
        fsim = np.where(np.isfinite(fsim), fsim, np.inf)

        def sort():
            order = np.argsort(fsim, axis=1, kind="stable")
            np.copyto(fsim, np.take_along_axis(fsim, order, axis=1))
            np.copyto(sim, np.take_along_axis(sim, order[:, :, None], axis=1))

        sort()
        live = np.ones(s_n, dtype=bool)
        # scipy counts the initial simplex as iteration 1 (``iterations = 1;
        # while iterations < maxiter``): num_steps = maxiter, as the reference
        # passes it (nelder_mead.py:84)
        for it in range(max(int(num_steps) - 1, 0)):
            conv = (np.abs(sim[:, 1:] - sim[:, :1]).max(axis=(1, 2)) <= xatol) & \
                   (np.abs(fsim[:, :1] - fsim[:, 1:]).max(axis=1) <= fatol)
            live &= ~conv
            if not live.any():
                break
            xbar = sim[:, :-1].sum(axis=1) / d
            worst = sim[:, -1]
            clip = lambda x: np.clip(x, self.low, self.high)    # noqa: E731
            xr = clip((1 + rho) * xbar - rho * worst)
            xe = clip((1 + rho * chi) * xbar - rho * chi * worst)
            xc = clip((1 + psi * rho) * xbar - psi * rho * worst)
            xcc = clip((1 - psi) * xbar + psi * worst)
            f = self._score(np.concatenate((xr, xe, xc, xcc)))  # one batched step
            f = np.where(np.isfinite(f), f, np.inf).reshape(4, s_n)
            fxr, fxe, fxc, fxcc = f
            shrink = np.zeros(s_n, dtype=bool)
            for s in np.nonzero(live)[0]:
                if fxr[s] < fsim[s, 0]:
                    if fxe[s] < fxr[s]:
                        sim[s, -1], fsim[s, -1] = xe[s], fxe[s]
                    else:
                        sim[s, -1], fsim[s, -1] = xr[s], fxr[s]
                elif fxr[s] < fsim[s, -2]:
                    sim[s, -1], fsim[s, -1] = xr[s], fxr[s]
                elif fxr[s] < fsim[s, -1]:
                    if fxc[s] <= fxr[s]:
                        sim[s, -1], fsim[s, -1] = xc[s], fxc[s]
                    else:
                        shrink[s] = True
                else:
                    if fxcc[s] < fsim[s, -1]:
                        sim[s, -1], fsim[s, -1] = xcc[s], fxcc[s]
                    else:
                        shrink[s] = True
            if shrink.any():
                idx = np.nonzero(shrink)[0]
                sim[idx, 1:] = clip(sim[idx, :1] + sigma * (sim[idx, 1:] - sim[idx, :1]))
                fs = self._score(sim[idx, 1:].reshape(-1, d)).reshape(len(idx), d)
                fsim[idx, 1:] = np.where(np.isfinite(fs), fs, np.inf)
            sort()
            best = float(fsim[:, 0].min())
            if self.verbose >= 2 and it % 10 == 0:
                print("NelderMead it %d best %.6g" % (it, best))
            if termination_fn is not None and termination_fn(best):
                break
        s_best = int(np.argmin(fsim[:, 0]))
        return sim[s_best, 0].copy(), float(fsim[s_best, 0])
