"""The policy-driven rollout (BASELINE.json configs[4]; the loop the reference
intends in policies/kshoot.py:65-85 / algorithm.py:564-620): the step kernel's
bfloat16 copy of the new state (mt_step_obs16), the fused output layer of the
actor (mt_actor_head) and the rollout that uses both."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

PARAMS = dict(dx=50, dt=0.5, rho_max=1, v_max=30, tau=9, lam=1, boundary_conditions="loop")


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (no CPU fallback exists)")


def make_obs(b, n, seed=3, dtype=np.float64):
    rng = np.random.default_rng(seed)
    rho = rng.uniform(0.05, 0.9, (b, n))
    v = 30 * (1 - rho) + rng.normal(0, 1, (b, n))
    return np.concatenate((rho, v), axis=1).astype(dtype)


@pytest.mark.parametrize("kind,dtype,b,n,extra", [
    ("arz", np.float64, 300, 1000, {}),            # pipelined kernel, 4 cells per thread
    ("arz", np.float64, 37, 250, {}),
    ("arz", np.float64, 16, 30, {}),               # two cells per thread
    ("lwr", np.float64, 64, 1000, {}),
    ("arz", np.float32, 64, 1000, {}),             # ring kernel: separate conversion pass
    ("arz", np.float32, 64, 1000, {"fast_math": True}),
    ("arz", np.float64, 5, 33, {}),                # scalar kernel
    ("arz", np.float64, 2, 40000, {}),             # long road: segment kernel
])
def test_step_obs16_is_the_bfloat16_of_the_new_state(kind, dtype, b, n, extra):
    _need_gpu()
    from mbrl_traffic_b200.engine import Engine
    name = np.dtype(dtype).name
    eng = Engine(kind, name, b, n)
    p = dict(PARAMS)
    if kind == "lwr":
        p.pop("tau")
    eng.set_params(**p, **extra)
    x = torch.as_tensor(make_obs(b, n, dtype=dtype), device="cuda:0")
    plain, out = torch.empty_like(x), torch.empty_like(x)
    o16 = torch.zeros(b, 2 * n, dtype=torch.bfloat16, device="cuda:0")
    rew = torch.empty(b, dtype=x.dtype, device="cuda:0")
    act = torch.full((b,), 27.0, dtype=x.dtype, device="cuda:0")
    eng.step(x, plain, action=act)
    eng.step(x, out, action=act, reward=rew, obs16=o16)
    torch.cuda.synchronize()
    assert torch.equal(out, plain)                       # the state itself is unchanged
    assert torch.equal(o16, out.to(torch.bfloat16))
    eng.close()


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_actor_head_matches_the_torch_layer(dtype):
    _need_gpu()
    from mbrl_traffic_b200.engine import actor_head
    g = torch.Generator(device="cuda:0").manual_seed(0)
    b, h = 1000, 256
    hid = torch.randn(b, h, generator=g, device="cuda:0").relu().to(torch.bfloat16)
    w = (0.1 * torch.randn(1, h, generator=g, device="cuda:0")).to(torch.bfloat16)
    bias = torch.tensor([0.05], device="cuda:0").to(torch.bfloat16)
    out = torch.empty(b, dtype=dtype, device="cuda:0")
    actor_head(hid, w, bias, out, 15.0)
    z = (hid.float() @ w.float().t()).squeeze(1) + bias.float()
    ref = 15.0 * (torch.tanh(z.to(torch.bfloat16).float()) + 1.0)
    torch.testing.assert_close(out.float(), ref, rtol=0, atol=0.15)   # one bf16 ulp of z near 0
    assert float((out.float() - ref).abs().mean()) < 0.02


def test_fused_policy_rollout_graph_equals_eager_and_tracks_the_plain_path():
    _need_gpu()
    from mbrl_traffic_b200 import ARZModel, ARZ_MODEL_PARAMS
    from mbrl_traffic_b200.rollout import BatchedMacroEnv, FeedForwardActor, PolicyRollout
    b, n, horizon = 512, 1000, 12
    p = dict(ARZ_MODEL_PARAMS)
    p.pop("optimizer_cls")
    model = ARZModel(None, None, None, None, 0, optimizer_cls=None,
                     action_is_speed_limit=True, **p)
    obs = torch.as_tensor(make_obs(b, n, seed=9), device="cuda:0")
    torch.manual_seed(1)
    actor = FeedForwardActor(2 * n).to("cuda:0").to(torch.bfloat16)
    env = BatchedMacroEnv(model, b, n, horizon=horizon, obs_bf16=True)
    runner = PolicyRollout(env, actor)
    assert runner.fused
    total = runner.run(obs).clone()
    final = env.obs.clone()
    # the same loop, eagerly, out of the same pieces
    env2 = BatchedMacroEnv(model, b, n, horizon=horizon, obs_bf16=True)
    env2.reset(obs)
    act = torch.empty(b, dtype=torch.float64, device="cuda:0")
    tot2 = torch.zeros(b, dtype=torch.float64, device="cuda:0")
    with torch.no_grad():
        for _ in range(horizon):
            actor.act_bf16(env2.obs16, act)
            _, r, _, _ = env2.step(act)
            tot2 += r
    assert torch.equal(final, env2.obs) and torch.equal(total, tot2)
    assert torch.equal(env2.obs16, env2.obs.to(torch.bfloat16))
    # first action: fused layers vs the plain module on the same bf16 observation
    with torch.no_grad():
        a_plain = actor(obs.to(torch.bfloat16)).float()
        a_fused = actor.act_bf16(obs.to(torch.bfloat16), act).float()
    assert float((a_plain - a_fused).abs().max()) < 0.5     # of a [0, 30] range, bf16 layers
    model.close()


@pytest.mark.parametrize("case", ["rm1", "one", "two", "half", "mixed"])
@pytest.mark.parametrize("b,n", [(300, 1000), (37, 250)])
def test_per_road_tables_with_a_shared_exponent_kind(case, b, n):
    """mt_set_params reads back whether all roads share an exponent kind (and
    whether every rho_max is 1) and the step then runs with the kind as a
    compile-time constant; every variant -- and the general kernel for mixed
    kinds -- is bit-equal to the oracle, single step, with rewards, and as a
    fused rollout."""
    _need_gpu()
    from mbrl_traffic_b200.engine import Engine
    from oracle import np_oracle as o
    rng = np.random.default_rng(len(case) * 1000 + n)
    lam = {"rm1": np.ones(b), "one": np.ones(b), "two": np.full(b, 2.0), "half": np.full(b, 0.5),
           "mixed": rng.choice([1.0, 2.0, 0.5], b)}[case]
    rho_max = np.ones(b) if case == "rm1" else rng.choice([1.0, 0.5, 0.2], b)
    v_max = rng.uniform(15, 30, b)
    tau = rng.uniform(5, 20, b)
    rho = rho_max[:, None] * rng.uniform(0.05, 0.9, (b, n))
    v = v_max[:, None] * (1 - rho / rho_max[:, None]) ** lam[:, None] + rng.normal(0, 1, (b, n))
    obs = np.concatenate((rho, v), axis=1)
    dev = "cuda:0"
    eng = Engine("arz", "float64", b, n)
    eng.set_params(dx=50, dt=0.5, rho_max=torch.as_tensor(rho_max, device=dev),
                   v_max=torch.as_tensor(v_max, device=dev), lam=torch.as_tensor(lam, device=dev),
                   tau=torch.as_tensor(tau, device=dev), boundary_conditions="loop")
    kw = dict(dx=50, dt=0.5, rho_max=rho_max, v_max=v_max, lam=lam, tau=tau,
              rho_crit=rho_max / 2, boundary_conditions="loop")
    x = torch.as_tensor(obs, device=dev)
    y = torch.empty_like(x)
    rew = torch.empty(b, dtype=torch.float64, device=dev)
    eng.step(x, y, reward=rew)
    ref = o.arz_step(obs, **kw)
    assert np.array_equal(y.cpu().numpy(), ref)
    np.testing.assert_allclose(rew.cpu().numpy(), o.mean_speed_reward(ref), rtol=1e-12)
    ref4 = obs
    for _ in range(4):
        ref4 = o.arz_step(ref4, **kw)
    out4 = eng.rollout(x.clone(), y, 4)
    assert np.array_equal(out4.cpu().numpy(), ref4)
    eng.close()


def test_population_loss_with_general_exponents_stays_within_the_promise():
    """float64 population scoring evaluates x**lam with the engine's own
    exp2(lam log2 x) (5e-15 relative): every candidate's loss equals the
    oracle's (NumPy pow) to 1e-12, far inside the 1e-13-per-cell promise for
    exponents outside {1, 2, 0.5}."""
    _need_gpu()
    from mbrl_traffic_b200 import ARZModel, ARZ_MODEL_PARAMS
    from oracle import np_oracle as o
    b, n, p = 64, 30, 16
    rng = np.random.default_rng(71)
    rho = 0.2 * rng.uniform(0.05, 0.9, (b, n))
    obs = np.concatenate((rho, 30 * (1 - rho / 0.2) + rng.normal(0, 1, (b, n))), axis=1)
    rho2 = 0.2 * rng.uniform(0.05, 0.9, (b, n))
    nxt = np.concatenate((rho2, 30 * (1 - rho2 / 0.2) + rng.normal(0, 1, (b, n))), axis=1)
    params = dict(ARZ_MODEL_PARAMS)
    params.pop("optimizer_cls")
    m = ARZModel(None, None, None, None, 0, optimizer_cls=None, **params)
    cand = np.column_stack([rng.uniform(0.55, 1.0, p), rng.uniform(5, 30, p),
                            rng.uniform(1, 25, p), rng.uniform(0.6, 3.5, p)])
    losses = m.population_loss(cand, obs, None, nxt)
    for i in range(p):
        ref = o.density_mse(nxt, o.arz_step(obs, rho_max=cand[i, 0], v_max=cand[i, 1],
                                            tau=cand[i, 2], lam=cand[i, 3], rho_crit=0.5))
        assert abs(losses[i] - ref) <= 1e-12 * abs(ref), (i, losses[i], ref)
    m.close()
