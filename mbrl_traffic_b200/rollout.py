"""Model-based rollouts on the batched engine.

The reference's intended MBRL loop -- a policy proposing speed limits, the
macroscopic model predicting the outcome, rewards summed over a horizon -- is
a stub there (``policies/kshoot.py:65-85`` raises ``NotImplementedError``).
This module is the thin PyTorch side of that loop; every model step is one
``mt_step`` / ``mt_rollout`` call and no tensor leaves the GPU inside a
rollout.

Conventions follow the reference's VSL environments
(``mbrl_traffic/envs/vsl.py``): the action is a speed limit in
``[0, v_max_max]`` (``vsl.py:92-100``), one per road; the reward is the mean
speed of the vehicles (``vsl.py:115-118``), whose macroscopic analogue is the
density-weighted mean speed ``sum(rho v) / sum(rho)`` that the step kernel
reduces per road.
"""
import torch
from torch import nn

from mbrl_traffic_b200.engine import dtype_name


class BatchedMacroEnv(object):
    """Gym-style facade over a batch of macroscopic roads on one GPU.

    ``model`` is an ``LWRModel`` / ``ARZModel`` / ``NonLocalModel``; its
    parameters define the dynamics.  Observations are ``(B, 2N)`` device
    tensors ``[rho | v]``.
    """

    def __init__(self, model, n_envs, n_cells, horizon=500, dtype=torch.float64,
                 device=None, obs_bf16=False):
        """``obs_bf16``: every step also leaves a bfloat16 copy of the new
        state in ``self.obs16`` (``mt_step_obs16``) for a policy network."""
        self.model = model
        self.n_envs, self.n_cells, self.horizon = n_envs, n_cells, horizon
        self.dtype = dtype
        self.device = torch.device(device if device is not None
                                   else "cuda:%d" % model.device)
        self.engine = model._get_engine(n_envs, n_cells, dtype_name(dtype), host=False)
        self._bufs = [torch.empty(n_envs, 2 * n_cells, dtype=dtype, device=self.device)
                      for _ in range(2)]
        self._cur = 0
        self.reward = torch.empty(n_envs, dtype=dtype, device=self.device)
        self.obs16 = torch.empty(n_envs, 2 * n_cells, dtype=torch.bfloat16,
                                 device=self.device) if obs_bf16 else None
        self.t = 0

    @property
    def obs(self):
        return self._bufs[self._cur]

    def reset(self, obs0):
        """Start an episode from ``obs0`` (copied)."""
        self._cur = 0
        self._bufs[0].copy_(obs0)
        if self.obs16 is not None:
            self.obs16.copy_(obs0)
        self.t = 0
        return self.obs

    def step(self, action=None):
        """Advance every road by one step under speed limits ``action``
        (``(B,)`` device tensor in the env dtype, or None to keep ``v_max``).
        Returns ``(obs, reward, done, info)``; ``reward`` is a device tensor
        that the next call overwrites."""
        self.model._sync_params(
            (self.n_envs, self.n_cells, dtype_name(self.dtype), self.model.device, False),
            self.engine)
        src, dst = self._bufs[self._cur], self._bufs[1 - self._cur]
        self.engine.step(src, dst, action=action, reward=self.reward, obs16=self.obs16)
        self._cur = 1 - self._cur
        self.t += 1
        return self.obs, self.reward, self.t >= self.horizon, {}


def _diverged(self):
    """int32 ``(B,)`` device tensor: 1 for every road whose current state holds
    an inf or a NaN (``mt_nonfinite``).  An erratic controller can drive the
    reference scheme itself unstable; such roads are usually reset."""
    if getattr(self, "_flags", None) is None:
        self._flags = torch.empty(self.n_envs, dtype=torch.int32, device=self.device)
    self.engine.nonfinite(self.obs, self._flags)
    return self._flags


BatchedMacroEnv.diverged = _diverged


class FeedForwardActor(nn.Module):
    """``[2N -> 256 -> 256 -> 1]`` actor, the shape of the reference's SAC
    policy network (``policies/sac.py:303-322``, ``utils/train.py:152``); the
    output is squashed to a speed limit in ``[0, v_max_max]``."""

    def __init__(self, obs_dim, v_max_max=30.0, hidden=(256, 256)):
        super(FeedForwardActor, self).__init__()
        layers, last = [], obs_dim
        for h in hidden:
            layers += [nn.Linear(last, h), nn.ReLU()]
            last = h
        layers.append(nn.Linear(last, 1))
        self.net = nn.Sequential(*layers)
        self.v_max_max = float(v_max_max)

    def forward(self, obs):
        x = self.net(obs.to(self.net[0].weight.dtype))
        return (0.5 * self.v_max_max) * (torch.tanh(x.squeeze(-1)) + 1.0)

    def act_bf16(self, obs16, action_out, reward_prev=None, return_sum=None):
        """The same policy for a bfloat16 network and a bfloat16 observation
        (``BatchedMacroEnv(obs_bf16=True).obs16``), written into
        ``action_out`` (``(B,)`` in the engine dtype): hidden layers as GEMMs
        with bias + ReLU in their epilogue, the output layer with its tanh and
        scaling as one small kernel (``mt_actor_head``) -- no cast of the
        observation, no elementwise kernels in between."""
        from mbrl_traffic_b200.engine import actor_head
        x = obs16
        linears = [m for m in self.net if isinstance(m, nn.Linear)]
        for lin in linears[:-1]:
            x = torch._addmm_activation(lin.bias, x, lin.weight.t())   # relu(x W^T + b)
        last = linears[-1]
        return actor_head(x, last.weight, last.bias, action_out, 0.5 * self.v_max_max,
                          reward_prev=reward_prev, return_sum=return_sum)


class PolicyRollout(object):
    """``actor`` rolled through ``env``, captured ONCE as a CUDA graph of two
    consecutive (policy, step, reward) iterations -- one per ping-pong
    direction of the environment's state buffers -- and replayed for every
    rollout after that: the loop has no per-step launch cost and no capture
    cost after the first call (the caller of evaluate_model.py:388-403 rolls
    the same policy out again and again)."""

    def __init__(self, env, actor):
        self.env, self.actor = env, actor
        self.total = torch.zeros(env.n_envs, dtype=env.dtype, device=env.device)
        self.action = torch.empty(env.n_envs, dtype=env.dtype, device=env.device)
        #: the fused path: bfloat16 actor on the bfloat16 observation the step
        #: kernel leaves behind (BatchedMacroEnv(obs_bf16=True))
        self.fused = (env.obs16 is not None and hasattr(actor, "act_bf16")
                      and next(actor.parameters()).dtype == torch.bfloat16)
        self.stream = torch.cuda.Stream(device=env.device)
        self.graph = None

    def _one(self):
        env = self.env
        with torch.no_grad():
            if self.fused:
                # the return accumulates on the output layer's launch: it adds the
                # PREVIOUS step's reward (zero before the first step; the last
                # step's reward is added by run())
                act = self.actor.act_bf16(env.obs16, self.action, reward_prev=env.reward,
                                          return_sum=self.total)
            else:
                act = self.actor(env.obs).to(env.dtype).contiguous()
        _, r, _, _ = env.step(act)
        if not self.fused:
            self.total.add_(r)

    def _capture(self):
        env = self.env
        with torch.cuda.stream(self.stream):
            self._one()
            self._one()                         # warm-up both directions
            torch.cuda.synchronize(env.device)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=self.stream):
                self._one()
                self._one()

    def run(self, obs0, horizon=None):
        """Sum of rewards per road over ``horizon`` steps from ``obs0``
        (device tensor, overwritten by the next call)."""
        env = self.env
        horizon = env.horizon if horizon is None else horizon
        if self.graph is None:
            self.stream.wait_stream(torch.cuda.current_stream(env.device))
            with torch.cuda.stream(self.stream):
                env.reset(obs0)
            self._capture()
        self.stream.wait_stream(torch.cuda.current_stream(env.device))
        with torch.cuda.stream(self.stream):
            env.reset(obs0)
            self.total.zero_()
            if self.fused:
                env.reward.zero_()
            done = 0
            while done + 2 <= horizon:
                self.graph.replay()
                done += 2
            for _ in range(horizon - done):
                self._one()
            if self.fused:
                self.total.add_(env.reward)        # the last step's reward
        torch.cuda.current_stream(env.device).wait_stream(self.stream)
        return self.total


def policy_rollout(env, actor, obs0, horizon=None, use_graph=True):
    """Roll ``actor`` through ``env`` for ``horizon`` steps; returns the sum
    of rewards per road (device tensor).  With ``use_graph`` the loop runs as
    replays of a two-step CUDA graph (``PolicyRollout``; keep that object to
    amortise the capture over many rollouts)."""
    horizon = env.horizon if horizon is None else horizon
    if use_graph and horizon >= 4:
        return PolicyRollout(env, actor).run(obs0, horizon)
    env.reset(obs0)
    total = torch.zeros(env.n_envs, dtype=env.dtype, device=env.device)
    for _ in range(horizon):
        with torch.no_grad():
            act = actor(env.obs).to(env.dtype).contiguous()
        _, r, _, _ = env.step(act)
        total.add_(r)
    return total


class KShootPolicy(object):
    """Random-shooting MPC over the model (the reference's ``KShootPolicy``,
    ``policies/kshoot.py:7-85``, whose ``get_action`` is unimplemented;
    parameters ``num_particles`` / ``traj_length`` from
    ``utils/train.py:130-135``).

    For every road, ``num_particles`` speed-limit sequences of length
    ``traj_length`` are sampled uniformly in ``[0, v_max_max]``, all
    ``B x num_particles`` copies are rolled through the model in one
    ``mt_rollout`` call, and the first action of the best-return sequence is
    returned.
    """

    def __init__(self, model, n_envs, n_cells, num_particles=16, traj_length=10,
                 v_max_max=30.0, dtype=torch.float64, seed=0):
        if not model.action_is_speed_limit:
            raise ValueError("the model must be built with action_is_speed_limit=True")
        self.model = model
        self.n_envs, self.n_cells = n_envs, n_cells
        self.num_particles, self.traj_length = num_particles, traj_length
        self.v_max_max, self.dtype = float(v_max_max), dtype
        self.device = torch.device("cuda:%d" % model.device)
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(seed)
        total = n_envs * num_particles
        self.engine = model._get_engine(total, n_cells, dtype_name(dtype), host=False)
        self._a = torch.empty(total, 2 * n_cells, dtype=dtype, device=self.device)
        self._b = torch.empty_like(self._a)
        self._rew = torch.empty(traj_length, total, dtype=dtype, device=self.device)

    def sample_actions(self):
        """``(traj_length, B * num_particles)`` speed limits; particle p of
        road b sits at column ``b * num_particles + p``."""
        u = torch.rand(self.traj_length, self.n_envs * self.num_particles,
                       generator=self.gen, device=self.device, dtype=self.dtype)
        return u * self.v_max_max

    def get_action(self, obs, actions=None):
        """Best first action per road, ``(B,)``; also returns the returns of
        all particles ``(B, num_particles)``."""
        if actions is None:
            actions = self.sample_actions()
        # the model may have been refitted / reloaded since the last call
        self.model._sync_params(
            (self.n_envs * self.num_particles, self.n_cells, dtype_name(self.dtype),
             self.model.device, False), self.engine)
        k = self.num_particles
        # every road's state to its k particles: one broadcasting copy
        self._a.view(self.n_envs, k, -1).copy_(obs[:, None, :].expand(-1, k, -1))
        self.engine.rollout(self._a, self._b, self.traj_length,
                            actions=actions.contiguous(), rewards=self._rew)
        returns = self._rew.sum(dim=0).view(self.n_envs, k)
        best = returns.argmax(dim=1)
        first = actions[0].view(self.n_envs, k)
        return first.gather(1, best[:, None]).squeeze(1), returns
