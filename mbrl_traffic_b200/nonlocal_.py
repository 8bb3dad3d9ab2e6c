"""Non-local look-ahead LWR model on the B200 engine.

The reference names a ``NonLocalModel`` in its CLI and stubs
(``utils/train.py:351,353,643,685-689``; ``evaluate_model.py:15,281-282``)
but never implemented it.  This class follows the same ``Model`` contract as
``LWRModel`` and the scheme of the paper the reference cites at
``lwr.py:209-211`` (Friedrich, Kolb, Goettlich, NHM 13, 2018): drivers adapt
their speed to a weighted mean of the downstream equilibrium speeds over a
look-ahead distance ``eta``.  ``scheme="lxf"`` selects the other classical
variant instead: the speed of a weighted mean of the downstream DENSITIES with
a Lax-Friedrichs flux (Blandin & Goatin, Numer. Math. 132, 2016).

Rounding (there is no reference code to follow): the look-ahead sums take
their terms in ascending k, each joined by ONE fused multiply-add,
``w = fma(gamma_k, V, w)``; everything else is separately rounded like the
LWR / ARZ steps.  ``oracle/np_oracle.py`` evaluates exactly that.
"""
import json
import os
from copy import deepcopy

import numpy as np

from mbrl_traffic_b200._fv import FiniteVolumeModel
from mbrl_traffic_b200.params import LAM_MAX


def lookahead_weights(eta, dx, kernel="linear"):
    """gamma_k = integral of w_eta over [k dx, (k+1) dx], k < eta/dx.

    ``w_eta`` is non-increasing on [0, eta] with unit mass: ``linear``
    w(y) = 2 (eta - y) / eta**2, ``constant`` w(y) = 1 / eta.  ``eta`` must be
    a whole number of cells.  Evaluated in float64 (the engine converts).
    """
    n_eta = int(round(eta / dx))
    if n_eta < 1 or abs(n_eta * dx - eta) > 1e-9 * max(1.0, abs(eta)):
        raise ValueError("eta must be a positive multiple of dx")
    k = np.arange(n_eta, dtype=np.float64)
    lo, hi = k * dx, (k + 1) * dx
    if kernel == "linear":
        return (2.0 * eta * (hi - lo) - (hi * hi - lo * lo)) / (eta * eta)
    if kernel == "constant":
        return (hi - lo) / eta
    raise ValueError("Unknown look-ahead kernel: {}".format(kernel))


class NonLocalModel(FiniteVolumeModel):
    """Non-local LWR model (mean downstream velocity), batched on the GPU.

    Same constructor shape as ``LWRModel`` plus ``eta`` (look-ahead distance
    in metres) and ``kernel`` (``"linear"`` | ``"constant"``).  Boundary
    conditions: ``"loop"`` is a true periodic ring for this model,
    ``"extend_both"``, ``{"constant_both": (rho_l, rho_r)}`` and
    ``{"inflow_outflow": rho_l}`` (fixed inflow density, free outflow).
    ``scheme``: ``"godunov"`` (default) or ``"lxf"``; ``alpha`` is the LxF
    viscosity (None: each road's ``v_max``, the bound of ``|V|``).

    ``get_next_obs`` keeps the reference observation layout ``[rho | v]``;
    like ``LWRModel`` the input velocity is ignored and the output velocity is
    the equilibrium speed of the new density.
    """

    _engine_model = "nonlocal"
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
                 eta,
                 kernel,
                 boundary_conditions,
                 optimizer_cls,
                 optimizer_params=None,
                 scheme="godunov",
                 alpha=None,
                 **engine_kwargs):
        if scheme not in ("godunov", "lxf"):
            raise ValueError("Unknown scheme: {}".format(scheme))
        super(NonLocalModel, self).__init__(
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
        self.eta = eta
        self.kernel = kernel
        self.scheme = scheme
        self.alpha = alpha
        self.boundary_conditions = boundary_conditions
        self.optimizer_cls = optimizer_cls
        self.optimizer_params = optimizer_params or {}

        def fitness_fn(x):
            """Loss of ``[rho_max, v_max, lam]`` on one replay batch."""
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

    def weights(self):
        """The look-ahead weights gamma_k for the current eta / dx / kernel."""
        return lookahead_weights(self.eta, self.dx, self.kernel)

    def _max_n_eta(self):
        return 256

    def _engine_kwargs(self):
        return {"gamma": self.weights(), "scheme": self.scheme, "alpha": self.alpha}

    def cfl_number(self):
        """Godunov scheme: dt/dx * (gamma_0 |V'|_max rho_max + |V|_max); the
        scheme is monotone for values <= 1 (Friedrich et al. 2018, Thm 3.1;
        lam = 1).  LxF scheme: dt/dx * (alpha + gamma_0 |V'|_max rho_max) with
        the same bound (Blandin & Goatin 2016, condition (3.5) up to their
        factor conventions)."""
        g0 = float(self.weights()[0])
        v = float(np.max(self.v_max))
        if self.scheme == "lxf":
            alpha = v if self.alpha is None else float(self.alpha)
            return self.dt / self.dx * (alpha + g0 * v)
        return self.dt / self.dx * (g0 * v + v)

    def update(self):
        if self.optimizer is None:
            raise RuntimeError("no optimizer: pass a class as optimizer_cls")
        model_params, error = self.optimizer.solve()
        self.rho_max, self.v_max, self.lam = model_params[:3]
        return error

    _KEYS = ("dx", "dt", "rho_max", "rho_max_max", "v_max", "v_max_max",
             "stream_model", "lam", "eta", "kernel", "boundary_conditions")
    _OPTIONAL_KEYS = ("scheme", "alpha")

    def save(self, save_path):
        """``model.json`` with the LWR keys plus ``eta`` and ``kernel``."""
        with open(os.path.join(save_path, "model.json"), 'w') as f:
            json.dump({k: getattr(self, k) for k in self._KEYS + self._OPTIONAL_KEYS}, f,
                      sort_keys=True, indent=4)

    def load(self, load_path):
        with open(load_path) as f:
            data = json.load(f)
        for key in self._KEYS:
            setattr(self, key, data[key])
        for key in self._OPTIONAL_KEYS:
            if key in data:
                setattr(self, key, data[key])
