"""ARZ model on the B200 engine -- drop-in for
``mbrl_traffic/models/arz.py`` (class ``ARZModel``, arz.py:13-558)."""
import json
import os
from copy import deepcopy

from mbrl_traffic_b200._fv import FiniteVolumeModel
from mbrl_traffic_b200.params import LAM_MAX


class ARZModel(FiniteVolumeModel):
    """Aw-Rascle-Zhang traffic flow model, batched on the GPU.

    Constructor arguments, attribute names, ``get_next_obs`` semantics and the
    ``model.json`` schema are the reference's (arz.py:60-145, 205-221,
    526-558).  The step -- relative flow ``y = rho (v - V(rho))``, supply/demand
    flux, semi-implicit relaxation and the three boundary modes
    (arz.py:223-467) -- runs in ``mt::step_pipe_kernel<T, 1, ...>``
    (csrc/mt_pipe.cuh; per-cell arithmetic in csrc/mt_cell.cuh ``arz_*``); the relaxation's
    ``fsolve`` (arz.py:462-463) is replaced by the closed-form root of the same
    affine system (see oracle/np_oracle.py ``arz_step``).

    ``rho_crit`` is frozen at construction exactly like the reference
    (arz.py:199): assigning ``rho_max`` later does not move it.

    Extra keyword arguments (not in the reference): ``device``,
    ``action_is_speed_limit``, ``fast_math``, ``fsolve_abort`` (mimic the
    reference on roads with an exact-zero density, where its fsolve call gives
    up and the relative flow stays frozen for the road, arz.py:411, 460-463;
    off by default: the closed form gives NaN at the affected cells only).
    ``rho_max``, ``v_max``, ``tau`` and ``lam`` may also be arrays with one
    value per row of a batched ``obs``.
    """

    _engine_model = "arz"
    _param_names = ("rho_max", "v_max", "lam", "tau")
    _fit_names = ("rho_max", "v_max", "tau", "lam")      # arz.py:156

    def __init__(self,
                 sess,
                 ob_space,
                 ac_space,
                 replay_buffer,
                 verbose,
                 dx,
                 dt,
                 rho_max,
                 rho_max_max,
                 v_max,
                 v_max_max,
                 tau,
                 tau_max,
                 lam,
                 boundary_conditions,
                 optimizer_cls,
                 optimizer_params=None,
                 **engine_kwargs):
        super(ARZModel, self).__init__(
            sess=sess, ob_space=ob_space, ac_space=ac_space,
            replay_buffer=replay_buffer, verbose=verbose, **engine_kwargs)

        self.dx = dx
        self.dt = dt
        self.rho_max = rho_max
        self.rho_max_max = rho_max_max
        self.v_max = v_max
        self.v_max_max = v_max_max
        self.tau = tau
        self.tau_max = tau_max
        self.lam = lam
        self.boundary_conditions = boundary_conditions
        self.optimizer_cls = optimizer_cls
        self.optimizer_params = optimizer_params or {}

        def fitness_fn(x):
            """Loss of ``[rho_max, v_max, tau, lam]`` on one replay batch
            (arz.py:147-187), with the sample unpacked in the buffer's real
            order."""
            orig = (deepcopy(self.rho_max), deepcopy(self.v_max),
                    deepcopy(self.tau), deepcopy(self.lam))
            self.rho_max, self.v_max, self.tau, self.lam = \
                x[0], x[1], x[2], x[3]
            states, actions, _, next_states, _ = self.replay_buffer.sample()
            loss = self.compute_loss(states, actions, next_states)
            self.rho_max, self.v_max, self.tau, self.lam = orig
            return loss

        self._fitness_fn = fitness_fn
        self.optimizer = self._make_optimizer(
            optimizer_cls, [0, 0, 0, 0], [rho_max_max, v_max_max, tau_max, LAM_MAX], fitness_fn)

        # critical density defined by the Green-shield Model (arz.py:199)
        self.rho_crit = self.rho_max / 2

    def _engine_kwargs(self):
        return {"rho_crit": self.rho_crit}

    def update(self):
        """See parent class (arz.py:494-504).  The reference stores the fitted
        relaxation time into ``tau_max`` (arz.py:501, a typo); it is stored in
        ``tau`` here."""
        if self.optimizer is None:
            raise RuntimeError("no optimizer: pass a class as optimizer_cls")
        model_params, error = self.optimizer.solve()
        self.rho_max = model_params[0]
        self.v_max = model_params[1]
        self.tau = model_params[2]
        self.lam = model_params[3]
        return error

    def save(self, save_path):
        """See parent class (arz.py:526-542); same ten keys."""
        params = {
            "dx": self.dx,
            "dt": self.dt,
            "rho_max": self.rho_max,
            "rho_max_max": self.rho_max_max,
            "v_max": self.v_max,
            "v_max_max": self.v_max_max,
            "tau": self.tau,
            "tau_max": self.tau_max,
            "lam": self.lam,
            "boundary_conditions": self.boundary_conditions,
        }
        with open(os.path.join(save_path, "model.json"), 'w') as f:
            json.dump(params, f, sort_keys=True, indent=4)

    def load(self, load_path):
        """See parent class (arz.py:544-558); ``rho_crit`` is NOT refreshed,
        as in the reference."""
        with open(load_path) as f:
            data = json.load(f)
        for key in ("dx", "dt", "rho_max", "rho_max_max", "v_max",
                    "v_max_max", "tau", "tau_max", "lam",
                    "boundary_conditions"):
            setattr(self, key, data[key])
