"""GPU tests of stream order across back-to-back launches.  Step kernels are
launched with programmatic dependent launch (the next launch's CTAs set up and
prefetch towards L2 while the previous one drains; with
MT_FLAG_INDEPENDENT_INPUT they even start loading).  Every result must be what
a fully serialised sequence gives -- bit-equal to the oracle -- whatever is
enqueued on the stream in between: other engines, partial grids, fused
rollouts, foreign kernels, copies and memsets.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import c_oracle  # noqa: E402

B, N = 4096, 1000       # fills every CTA slot of a B200 (one road per tile)
PARAMS = dict(dx=50, dt=0.5, rho_max=1, v_max=30, tau=9, lam=1,
              boundary_conditions="loop")


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (no CPU fallback exists)")


def make_obs(seed, b=B, n=N):
    rng = np.random.default_rng(4321 + seed)
    rho = rng.uniform(0.05, 0.9, (b, n))
    v = 30.0 * (1.0 - rho) + rng.normal(0.0, 1.0, (b, n))
    return np.concatenate((rho, v), axis=1)


def ref_step(kind, obs, steps=1):
    out = obs
    for _ in range(steps):
        out = c_oracle.step(kind, out, **PARAMS)
    return out


def engine(kind="arz", b=B, n=N, mode="auto"):
    """mode: "auto" (the library checks the byte ranges itself), "vouched"
    (MT_FLAG_INDEPENDENT_INPUT) or "dependent" (MT_FLAG_DEPENDENT_INPUT)."""
    from mbrl_traffic_b200.engine import Engine
    eng = Engine(kind, "float64", b, n)
    eng.set_params(independent_input=(mode == "vouched"),
                   dependent_input=(mode == "dependent"), **PARAMS)
    return eng


@pytest.mark.parametrize("mode", ["auto", "vouched", "dependent"])
@pytest.mark.parametrize("kind", ["arz", "lwr"])
def test_pool_of_batches_eager_and_graph_bit_exact(kind, mode):
    """A pool of four full-GPU batches stepped round-robin, eagerly and from a
    replayed CUDA graph, with the early loads decided by the library, vouched
    for by the caller, or switched off: every output equals the oracle's."""
    _need_gpu()
    eng = engine(kind, mode=mode)
    obs = [make_obs(j) for j in range(4)]
    ins = [torch.as_tensor(x, device="cuda:0") for x in obs]
    outs = [torch.zeros_like(x) for x in ins]
    refs = [ref_step(kind, x) for x in obs]
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for k in range(40):
            eng.step(ins[k % 4], outs[k % 4], stream=s)
        s.synchronize()
        for j in range(4):
            assert np.array_equal(outs[j].cpu().numpy(), refs[j])
            outs[j].zero_()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for k in range(24):
                eng.step(ins[k % 4], outs[k % 4], stream=s)
        for _ in range(5):
            g.replay()
        s.synchronize()
    for j in range(4):
        assert np.array_equal(outs[j].cpu().numpy(), refs[j])
    eng.close()


def test_interleaved_chains_respect_older_launches():
    """Two chains A->B->A.. and C->D->C.. stepped alternately: each launch is
    independent of the one right before it but reads what the launch before
    THAT one wrote.  30 steps of each chain equal 30 oracle steps."""
    _need_gpu()
    eng = engine("arz")
    x0, y0 = make_obs(10), make_obs(11)
    a, c = torch.as_tensor(x0, device="cuda:0"), torch.as_tensor(y0, device="cuda:0")
    b, d = torch.empty_like(a), torch.empty_like(c)
    s = torch.cuda.Stream()
    steps = 30
    with torch.cuda.stream(s):
        for _ in range(steps):
            eng.step(a, b, stream=s)
            eng.step(c, d, stream=s)
            a, b = b, a
            c, d = d, c
        s.synchronize()
    assert np.array_equal(a.cpu().numpy(), ref_step("arz", x0, steps))
    assert np.array_equal(c.cpu().numpy(), ref_step("arz", y0, steps))
    eng.close()


def test_chained_steps_at_full_size():
    """input == the previous launch's output, 40 times, one road per tile on
    every CTA slot of the GPU."""
    _need_gpu()
    eng = engine("arz")
    x0 = make_obs(20)
    a = torch.as_tensor(x0, device="cuda:0")
    b = torch.empty_like(a)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(40):
            eng.step(a, b, stream=s)
            a, b = b, a
        s.synchronize()
    assert np.array_equal(a.cpu().numpy(), ref_step("arz", x0, 40))
    eng.close()


def test_foreign_work_between_steps_is_seen():
    """Other kernels and copies enqueued between two steps rewrite the next
    step's input: the step must read the rewritten data."""
    _need_gpu()
    eng = engine("arz")
    pool = [torch.as_tensor(make_obs(30 + j), device="cuda:0") for j in range(2)]
    outs = [torch.empty_like(pool[0]) for _ in range(2)]
    fresh = [make_obs(40 + j) for j in range(6)]
    fresh_dev = [torch.as_tensor(x, device="cuda:0") for x in fresh]
    fresh_pin = [torch.as_tensor(x).pin_memory() for x in fresh]
    refs = [ref_step("arz", x) for x in fresh]
    s = torch.cuda.Stream()
    got = []
    with torch.cuda.stream(s):
        for k in range(6):
            j = k % 2
            eng.step(pool[1 - j], outs[1 - j], stream=s)      # keeps the GPU busy
            if k % 3 == 0:
                pool[j].copy_(fresh_dev[k])                    # a kernel
            elif k % 3 == 1:
                pool[j].copy_(fresh_pin[k], non_blocking=True)  # a host-to-device copy
            else:
                pool[j].zero_()                                # a memset ...
                pool[j].add_(fresh_dev[k])                     # ... and a kernel
            eng.step(pool[j], outs[j], stream=s)
            got.append(outs[j].clone())
        s.synchronize()
    for k in range(6):
        assert np.array_equal(got[k].cpu().numpy(), refs[k]), k
    eng.close()


def test_mixed_engines_and_partial_grids_on_one_stream():
    """Launches of two engines (one filling the GPU, one not) and of the fused
    multi-step kernel share a stream; every result equals the oracle's."""
    _need_gpu()
    big, small = engine("arz"), engine("arz", b=300)
    xb = [make_obs(50 + j) for j in range(3)]
    xs = make_obs(60, b=300)
    ib = [torch.as_tensor(x, device="cuda:0") for x in xb]
    ob = [torch.empty_like(x) for x in ib]
    is_ = torch.as_tensor(xs, device="cuda:0")
    os_, tmp = torch.empty_like(is_), torch.empty_like(is_)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        big.step(ib[0], ob[0], stream=s)
        big.step(ib[1], ob[1], stream=s)
        small.step(is_, os_, stream=s)             # partial grid after overlapped launches
        big.step(ib[2], ob[2], stream=s)
        res = small.rollout(os_, tmp, 5, stream=s)  # fused kernel
        big.step(ob[0], ib[0], stream=s)           # reads the first launch's output
        s.synchronize()
    for j in range(3):
        assert np.array_equal(ob[j].cpu().numpy(), ref_step("arz", xb[j]))
    assert np.array_equal(ib[0].cpu().numpy(), ref_step("arz", xb[0], 2))
    assert np.array_equal(res.cpu().numpy(),
                          ref_step("arz", xs, 6))
    big.close()
    small.close()


def test_two_engines_share_the_stream_record():
    """Engine 2 writes the batch engine 1 reads next on the same stream."""
    _need_gpu()
    e1, e2 = engine("arz"), engine("arz")
    x0, y0 = make_obs(70), make_obs(71)
    a = torch.as_tensor(x0, device="cuda:0")
    c = torch.as_tensor(y0, device="cuda:0")
    b, d = torch.empty_like(a), torch.empty_like(a)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(10):
            e1.step(c, d, stream=s)      # independent filler
            e2.step(a, b, stream=s)      # engine 2: a -> b
            e1.step(b, a, stream=s)      # engine 1 reads what engine 2 just wrote
        s.synchronize()
    assert np.array_equal(a.cpu().numpy(), ref_step("arz", x0, 20))
    assert np.array_equal(d.cpu().numpy(), ref_step("arz", y0))
    e1.close()
    e2.close()
