"""Shared host logic of the finite-volume models (LWR / ARZ / non-local).

The subclasses only declare their constructor (the reference signatures) and
which attributes feed the engine.  Everything numerical happens in
``libmtraffic.so``; this module converts observations, caches one engine per
``(batch, cells, dtype)`` and re-derives the device parameter table whenever
a model attribute changed since the last call -- the reference mutates
attributes between steps (fitness_fn lwr.py:161-164, arz.py:169-173; load();
the tests at tests/fast_tests/test_models.py:100-162 assign them directly).
"""
import numpy as np

from mbrl_traffic_b200.base import Model
from mbrl_traffic_b200.engine import Engine, PinnedPool, dtype_name


def _is_torch(x):
    return hasattr(x, "data_ptr") and hasattr(x, "device")


class FiniteVolumeModel(Model):
    """Common machinery; see ``LWRModel`` / ``ARZModel`` / ``NonLocalModel``."""

    #: engine model name, set by subclasses
    _engine_model = None
    #: attributes that may be scalars or per-env arrays
    _param_names = ("rho_max", "v_max", "lam")

    def __init__(self, sess, ob_space, ac_space, replay_buffer, verbose,
                 device=0, action_is_speed_limit=False, fast_math=False,
                 fsolve_abort=False):
        super(FiniteVolumeModel, self).__init__(
            sess=sess, ob_space=ob_space, ac_space=ac_space,
            replay_buffer=replay_buffer, verbose=verbose)
        #: CUDA device ordinal of the engine
        self.device = device
        #: if True, ``action`` (one value per env) overrides ``v_max`` for the
        #: step -- the use the reference documents (lwr.py:98-101) but never
        #: implemented (it deletes the action, lwr.py:194, arz.py:207)
        self.action_is_speed_limit = action_is_speed_limit
        #: allow FMA contraction / approximate division (not bit-comparable)
        self.fast_math = fast_math
        #: ARZ only: mimic the reference on roads with an exact-zero density,
        #: where its fsolve call gives up and the relative flow stays frozen
        #: for the whole road (arz.py:411, 460-463; MT_FLAG_FSOLVE_ABORT)
        self.fsolve_abort = fsolve_abort
        self._engines = {}
        self._param_keys = {}
        self._pop_cache = {}
        self._pinned = None

    # ------------------------------------------------------------------ #
    # reference-compatible no-ops                                        #
    # ------------------------------------------------------------------ #
    def initialize(self):
        """See parent class (lwr.py:188-190, arz.py:201-203)."""
        pass

    def get_td_map(self):
        """See parent class (lwr.py:345-347, arz.py:522-524)."""
        return {}

    # ------------------------------------------------------------------ #
    # engine plumbing                                                    #
    # ------------------------------------------------------------------ #
    def _max_n_eta(self):
        return 0

    def _engine_kwargs(self):
        """Extra ``Engine.set_params`` keywords (tau, rho_crit, gamma...)."""
        return {}

    def _get_engine(self, n_envs, n_cells, dtype, host):
        key = (int(n_envs), int(n_cells), dtype, self.device, bool(host))
        eng = self._engines.get(key)
        if eng is None:
            eng = Engine(self._engine_model, dtype, n_envs, n_cells,
                         device=self.device, max_n_eta=self._max_n_eta(),
                         host_staging=host)
            self._engines[key] = eng
        self._sync_params(key, eng)
        return eng

    def _snapshot(self, value):
        """Hashable identity of a parameter value for change detection."""
        if _is_torch(value):
            return ("t", value.data_ptr(), value._version, tuple(value.shape))
        if isinstance(value, np.ndarray):
            return ("a", value.tobytes(), value.shape)
        if isinstance(value, (dict, list, tuple)):
            return ("r", repr(value))
        return ("s", value)

    def _param_values(self):
        d = dict(dx=self.dx, dt=self.dt,
                 boundary_conditions=self.boundary_conditions)
        for name in self._param_names:
            d[name] = getattr(self, name)
        d.update(self._engine_kwargs())
        return d

    def _sync_params(self, key, eng):
        values = self._param_values()
        snap = tuple((k, self._snapshot(v)) for k, v in sorted(values.items()))
        snap += (("fast", self.fast_math), ("abort", self.fsolve_abort))
        if self._param_keys.get(key) == snap:
            return
        kw = {}
        for name, value in values.items():
            if isinstance(value, np.ndarray):
                import torch
                value = torch.as_tensor(
                    np.ascontiguousarray(value, dtype=eng.dtype),
                    device="cuda:%d" % self.device)
            kw[name] = value
        stream = None if any(_is_torch(v) for v in kw.values()) else 0
        eng.set_params(fast_math=self.fast_math, fsolve_abort=self.fsolve_abort,
                       stream=stream, **kw)
        self._param_keys[key] = snap

    # ------------------------------------------------------------------ #
    # the operator                                                       #
    # ------------------------------------------------------------------ #
    def get_next_obs(self, obs, action):
        """Advance a batch of roads by one step.

        Same contract as the reference (base.py:58-74): ``obs`` is
        ``(2N,)`` or ``(batch, 2N)`` laid out ``[rho | v]`` per row; a new
        array of the same shape/type is returned and ``obs`` is not modified.
        ``action`` is discarded unless ``action_is_speed_limit`` is set.
        NumPy in -> NumPy out (host path, copies included); torch CUDA tensor
        in -> torch CUDA tensor out (device path, asynchronous).
        """
        return self.run(obs, action, n_steps=1)

    def run(self, obs, action=None, n_steps=1, reward_out=None):
        """``n_steps`` chained ``get_next_obs`` calls in one library call
        (the loop of experiments/evaluate_model.py:398-403)."""
        if not self.action_is_speed_limit:
            action = None
        if _is_torch(obs):
            return self._run_device(obs, action, n_steps, reward_out)
        return self._run_host(obs, action, n_steps, reward_out)

    def _run_host(self, obs, action, n_steps, reward_out):
        arr = np.asarray(obs)
        if arr.dtype not in (np.float32, np.float64):
            arr = arr.astype(np.float64)
        squeeze = arr.ndim == 1
        batch = np.ascontiguousarray(arr[None, :] if squeeze else arr)
        if batch.ndim != 2 or batch.shape[1] % 2:
            raise ValueError("obs must be (2N,) or (batch, 2N)")
        n_envs, n_cells = batch.shape[0], batch.shape[1] // 2
        eng = self._get_engine(n_envs, n_cells, batch.dtype.name, host=True)
        # a NEW array per call, as in the reference -- on a recycled page-locked
        # block, so that the download runs at the PCIe rate (engine.PinnedPool)
        if self._pinned is None:
            self._pinned = PinnedPool()
        out = self._pinned.take(batch.shape, batch.dtype)
        act = None
        if action is not None:
            act = np.ascontiguousarray(
                np.asarray(action, dtype=batch.dtype).reshape(n_envs))
        eng.step_host(batch, out, action=act, reward=reward_out,
                      n_steps=n_steps)
        return out[0] if squeeze else out

    def _run_device(self, obs, action, n_steps, reward_out):
        import torch
        squeeze = obs.dim() == 1
        batch = obs.unsqueeze(0) if squeeze else obs
        batch = batch.contiguous()
        if batch.dim() != 2 or batch.shape[1] % 2:
            raise ValueError("obs must be (2N,) or (batch, 2N)")
        n_envs, n_cells = batch.shape[0], batch.shape[1] // 2
        eng = self._get_engine(n_envs, n_cells, dtype_name(batch.dtype),
                               host=False)
        out = torch.empty_like(batch)
        if action is not None:
            action = action.to(batch.dtype).reshape(n_envs).contiguous()
        if reward_out is not None and reward_out.numel() != n_envs:
            raise ValueError("reward_out must hold one value per road (the "
                             "final step's reward), as on the host path")
        if n_steps == 1:
            eng.step(batch, out, action=action, reward=reward_out)
        else:
            # the same speed limits apply to every step (as on the host path);
            # reward_out receives the final step's rewards: that step is issued
            # on its own, the steps before it as one fused launch
            work = batch.clone()
            fused = n_steps if reward_out is None else n_steps - 1
            acts = None if action is None else \
                action[None, :].expand(fused, n_envs).contiguous()
            res = eng.rollout(work, out, fused, actions=acts)
            if reward_out is not None:
                other = out if res is work else work
                eng.step(res, other, action=action, reward=reward_out)
                res = other
            out = res
        return out[0] if squeeze else out

    def trajectory(self, obs, actions=None, n_steps=1):
        """The states of ``n_steps`` chained steps, ``(n_steps + 1, batch, 2N)``
        (``(n_steps + 1, 2N)`` for one road), first entry = ``obs``: the loop of
        experiments/evaluate_model.py:388-414, which keeps every frame for its
        CSV / heat map.  The frames are written where they stay -- step k reads
        slice k of the trajectory buffer on the device and writes slice k + 1
        -- so a NumPy caller pays ONE upload and ONE download instead of a
        PCIe round trip per step.  ``actions``: ``(n_steps, batch)`` speed
        limits (used only with ``action_is_speed_limit``)."""
        import torch
        dev = "cuda:%d" % self.device
        was_torch = _is_torch(obs)
        if was_torch:
            first = obs
        else:
            arr = np.asarray(obs)
            if arr.dtype not in (np.float32, np.float64):
                arr = arr.astype(np.float64)
            first = torch.as_tensor(np.ascontiguousarray(arr), device=dev)
        squeeze = first.dim() == 1
        first = (first.unsqueeze(0) if squeeze else first).contiguous()
        if first.dim() != 2 or first.shape[1] % 2:
            raise ValueError("obs must be (2N,) or (batch, 2N)")
        n_envs, n_cells = first.shape[0], first.shape[1] // 2
        eng = self._get_engine(n_envs, n_cells, dtype_name(first.dtype), host=False)
        traj = torch.empty((n_steps + 1,) + tuple(first.shape), dtype=first.dtype,
                           device=first.device)
        traj[0].copy_(first)
        acts = None
        if actions is not None and self.action_is_speed_limit:
            acts = torch.as_tensor(actions, device=first.device).to(first.dtype) \
                .reshape(n_steps, n_envs).contiguous()
        for k in range(n_steps):
            eng.step(traj[k], traj[k + 1], action=None if acts is None else acts[k])
        out = traj[:, 0] if squeeze else traj
        return out if was_torch else out.cpu().numpy()

    # ------------------------------------------------------------------ #
    # loss                                                               #
    # ------------------------------------------------------------------ #
    def compute_loss(self, states, actions, next_states):
        """Density-half MSE between ``next_states`` and the model's
        prediction (lwr.py:329-343, arz.py:506-520), one batched step instead
        of the reference's per-row Python loop."""
        import torch
        dev = "cuda:%d" % self.device
        if _is_torch(states):
            s = states.contiguous()
            t = next_states.to(s.dtype).contiguous()
        else:
            arr = np.asarray(states)
            if arr.dtype not in (np.float32, np.float64):
                arr = arr.astype(np.float64)
            s = torch.as_tensor(np.ascontiguousarray(arr), device=dev)
            t = torch.as_tensor(
                np.ascontiguousarray(np.asarray(next_states, dtype=arr.dtype)),
                device=dev)
        act = None
        if self.action_is_speed_limit and actions is not None:
            act = torch.as_tensor(np.asarray(actions), device=dev) \
                if not _is_torch(actions) else actions
            act = act.to(s.dtype).reshape(s.shape[0]).contiguous()
        n_envs, n_cells = s.shape[0], s.shape[1] // 2
        eng = self._get_engine(n_envs, n_cells, dtype_name(s.dtype),
                               host=False)
        pred = torch.empty_like(s)
        eng.step(s, pred, action=act)
        sq = torch.empty(n_envs, dtype=s.dtype, device=s.device)
        eng.density_sqerr(pred, t, sq)
        return float(sq.sum().item() / (n_envs * n_cells))

    #: order of the fitted parameters in a candidate vector
    _fit_names = ("rho_max", "v_max", "lam")

    def population_loss(self, candidates, states, actions, next_states):
        """``compute_loss`` for a whole population of parameter vectors in one
        batched step: candidate ``p`` becomes the per-road parameters of its
        own copy of the ``B`` rows (``P*B`` roads in one ``mt_step``).

        ``candidates`` is ``(P, D)`` in the order of ``fitness_fn``'s ``x``
        (lwr.py:149, arz.py:156).  Returns a NumPy array of ``P`` losses,
        each equal to ``compute_loss`` with that candidate's parameters.
        """
        import torch
        dev = "cuda:%d" % self.device
        cand = np.asarray(candidates, dtype=np.float64)
        arr = np.asarray(states.cpu() if _is_torch(states) else states)
        if arr.dtype not in (np.float32, np.float64):
            arr = arr.astype(np.float64)
        tgt = np.asarray(next_states.cpu() if _is_torch(next_states) else next_states,
                         dtype=arr.dtype)
        p, b, n = cand.shape[0], arr.shape[0], arr.shape[1] // 2
        tdt = torch.float64 if arr.dtype == np.float64 else torch.float32
        # one engine and one set of device buffers per population shape, kept
        # across generations (an optimizer calls this once per generation)
        key = (p, b, n, arr.dtype.name, self.device)
        ws = self._pop_cache.get(key)
        if ws is None:
            eng = Engine(self._engine_model, arr.dtype.name, p * b, n, device=self.device,
                         max_n_eta=self._max_n_eta())
            ws = dict(eng=eng,
                      s=torch.empty(p * b, 2 * n, dtype=tdt, device=dev),
                      t=torch.empty(p * b, 2 * n, dtype=tdt, device=dev),
                      pred=torch.empty(p * b, 2 * n, dtype=tdt, device=dev),
                      sq=torch.empty(p * b, dtype=tdt, device=dev))
            self._pop_cache[key] = ws
        eng = ws["eng"]
        # every candidate gets its own copy of the rows: one broadcasting copy
        ws["s"].view(p, b, 2 * n).copy_(
            torch.as_tensor(np.ascontiguousarray(arr), device=dev)[None])
        ws["t"].view(p, b, 2 * n).copy_(
            torch.as_tensor(np.ascontiguousarray(tgt), device=dev)[None])
        kw = dict(dx=self.dx, dt=self.dt,
                  boundary_conditions=self.boundary_conditions)
        for name in self._param_names:
            kw[name] = getattr(self, name)
        kw.update(self._engine_kwargs())
        for j, name in enumerate(self._fit_names):
            kw[name] = torch.as_tensor(cand[:, j], device=dev).to(tdt) \
                .repeat_interleave(b).contiguous()
        if "gamma" in kw and isinstance(kw["gamma"], np.ndarray):
            kw["gamma"] = torch.as_tensor(kw["gamma"], device=dev).to(tdt)
        # float64 scoring: a general exponent goes through the engine's own
        # exp2(lam log2 x) (<= 5e-15 relative on the model's domain, well inside
        # the 1e-13 promised for exponents outside {1, 2, 0.5}; +55 % candidates
        # per second) -- in float64 the flag changes nothing else.  float32 keeps
        # the model's own setting (there the flag also contracts FMAs).
        fast = self.fast_math or tdt == torch.float64
        eng.set_params(fast_math=fast, fsolve_abort=self.fsolve_abort, **kw)
        eng.step(ws["s"], ws["pred"])
        eng.density_sqerr(ws["pred"], ws["t"], ws["sq"])
        return (ws["sq"].view(p, b).sum(dim=1) / (b * n)).double().cpu().numpy()

    def _make_optimizer(self, optimizer_cls, param_low, param_high, fitness_fn):
        """Build the optimizer the way the reference does (lwr.py:180-186);
        population optimizers of this package also get the batched scorer."""
        if not callable(optimizer_cls):
            return None
        kw = dict(self.optimizer_params)
        if getattr(optimizer_cls, "accepts_batch_fitness", False):
            def batch_fitness(population):
                states, actions, _, next_states, _ = self.replay_buffer.sample()
                return self.population_loss(population, states, actions, next_states)
            kw.setdefault("batch_fitness_fn", batch_fitness)
        return optimizer_cls(param_low=param_low, param_high=param_high,
                             fitness_fn=fitness_fn, verbose=self.verbose, **kw)

    # ------------------------------------------------------------------ #
    def close(self):
        """Release the engines (device memory)."""
        for eng in self._engines.values():
            eng.close()
        for ws in self._pop_cache.values():
            ws["eng"].close()
        if self._pinned is not None:
            self._pinned.drain()
        self._engines = {}
        self._param_keys = {}
        self._pop_cache = {}
