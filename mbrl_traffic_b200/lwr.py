"""LWR model on the B200 engine -- drop-in for
``mbrl_traffic/models/lwr.py`` (class ``LWRModel``, lwr.py:14-379)."""
import json
import os
from copy import deepcopy

from mbrl_traffic_b200._fv import FiniteVolumeModel
from mbrl_traffic_b200.params import LAM_MAX


class LWRModel(FiniteVolumeModel):
    """Lighthill-Whitham-Richards traffic flow model, batched on the GPU.

    Constructor arguments, attribute names, ``get_next_obs`` semantics and the
    ``model.json`` schema are the reference's (lwr.py:56-138, 192-204,
    349-379).  The step itself -- Greenshields speed, demand/supply Godunov
    flux with the reference's per-cell ``q_max``, frozen cell 0 and the three
    boundary modes (lwr.py:206-316) -- runs in ``mt::step_pipe_kernel<T, 0, ...>``
    (csrc/mt_pipe.cuh; per-cell arithmetic in csrc/mt_cell.cuh ``lwr_*``).

    Extra keyword arguments (not in the reference): ``device``,
    ``action_is_speed_limit``, ``fast_math``.  ``rho_max``, ``v_max`` and
    ``lam`` may also be arrays with one value per row of a batched ``obs``.
    """

    _engine_model = "lwr"
    _param_names = ("rho_max", "v_max", "lam")

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
                 stream_model,
                 lam,
                 boundary_conditions,
                 optimizer_cls,
                 optimizer_params=None,
                 **engine_kwargs):
        super(LWRModel, self).__init__(
            sess=sess, ob_space=ob_space, ac_space=ac_space,
            replay_buffer=replay_buffer, verbose=verbose, **engine_kwargs)

        self.dx = dx
        self.dt = dt
        self.rho_max = rho_max
        self.rho_max_max = rho_max_max
        self.v_max = v_max
        self.v_max_max = v_max_max
        self.stream_model = stream_model
        self.lam = lam
        self.boundary_conditions = boundary_conditions
        self.optimizer_cls = optimizer_cls
        self.optimizer_params = optimizer_params or {}

        def fitness_fn(x):
            """Loss of the parameters ``[rho_max, v_max, lam]`` on one replay
            batch (lwr.py:140-177).  The reference unpacks the sample in the
            wrong order; here ``obs1`` is the target."""
            orig = (deepcopy(self.rho_max), deepcopy(self.v_max),
                    deepcopy(self.lam))
            self.rho_max, self.v_max, self.lam = x[0], x[1], x[2]
            states, actions, _, next_states, _ = self.replay_buffer.sample()
            loss = self.compute_loss(states, actions, next_states)
            self.rho_max, self.v_max, self.lam = orig
            return loss

        self._fitness_fn = fitness_fn
        self.optimizer = self._make_optimizer(
            optimizer_cls, [0, 0, 0], [rho_max_max, v_max_max, LAM_MAX], fitness_fn)

    def update(self):
        """See parent class (lwr.py:318-327)."""
        if self.optimizer is None:
            raise RuntimeError("no optimizer: pass a class as optimizer_cls")
        model_params, error = self.optimizer.solve()
        self.rho_max = model_params[0]
        self.v_max = model_params[1]
        self.lam = model_params[2]
        return error

    def save(self, save_path):
        """See parent class (lwr.py:349-364); same nine keys."""
        params = {
            "dx": self.dx,
            "dt": self.dt,
            "rho_max": self.rho_max,
            "rho_max_max": self.rho_max_max,
            "v_max": self.v_max,
            "v_max_max": self.v_max_max,
            "stream_model": self.stream_model,
            "lam": self.lam,
            "boundary_conditions": self.boundary_conditions,
        }
        with open(os.path.join(save_path, "model.json"), 'w') as f:
            json.dump(params, f, sort_keys=True, indent=4)

    def load(self, load_path):
        """See parent class (lwr.py:366-379)."""
        with open(load_path) as f:
            data = json.load(f)
        for key in ("dx", "dt", "rho_max", "rho_max_max", "v_max",
                    "v_max_max", "stream_model", "lam",
                    "boundary_conditions"):
            setattr(self, key, data[key])
