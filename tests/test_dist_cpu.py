"""Host-side logic of the multi-GPU paths on CPU: road sharding and the
domain-decomposed long road with halo exchange over a world_size-2 ``gloo``
group.  The local stepper here is the NumPy oracle (the checker), so the test
exercises exactly the partitioning / ghost-cell / exchange logic that the GPU
driver uses around ``mt_step``."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mbrl_traffic_b200.dist import DomainDecomposedRoad, shard_range
from oracle import np_oracle as o


def test_shard_range_covers_everything_once():
    for n, w in [(4096, 8), (10, 3), (7, 7), (5, 8), (16384, 4)]:
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for a, b in zip(spans, spans[1:]):
            assert a[1] == b[0]
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def _oracle_stepper(kind, bc):
    fn = o.arz_step if kind == "arz" else o.lwr_step

    def step(local, bc_left, bc_right):
        # ghost edges are overwritten by the exchange; any rule will do there,
        # so the reference's own mode is applied to the whole block
        return fn(local, boundary_conditions=bc)
    return step


def _make_road(n, seed=0):
    rng = np.random.default_rng(seed)
    rho = rng.uniform(0.05, 0.9, n)
    v = 30 * (1 - rho) + rng.normal(0, 1, n)
    return np.concatenate((rho, v))


@pytest.mark.parametrize("kind", ["lwr", "arz"])
@pytest.mark.parametrize("world,halo", [(2, 1), (3, 4), (4, 8)])
def test_decomposed_road_single_process_emulation(kind, world, halo):
    """All ranks stepped in one process, halos delivered by hand."""
    n, steps = 240, 24
    bc = "extend_both"
    road = _make_road(n, seed=world)
    ref = road
    fn = o.arz_step if kind == "arz" else o.lwr_step
    for _ in range(steps):
        ref = fn(ref, boundary_conditions=bc)
    ranks = [DomainDecomposedRoad(n, r, world, halo, _oracle_stepper(kind, bc),
                                  exchange=lambda *_: (None, None))
             for r in range(world)]
    local = [d.scatter(road) for d in ranks]
    for _ in range(steps):
        if ranks[0].needs_refresh():
            edges = [d.edge_cells(x) for d, x in zip(ranks, local)]
            for r, d in enumerate(ranks):
                recv_l = edges[r - 1][1] if d.has_left else None
                recv_r = edges[r + 1][0] if d.has_right else None
                d.set_ghosts(local[r], recv_l, recv_r)
        local = [d.stepper(x, None, None) for d, x in zip(ranks, local)]
        for d in ranks:
            d._since_exchange += 1
    rho = np.concatenate([d.own(x)[0] for d, x in zip(ranks, local)])
    v = np.concatenate([d.own(x)[1] for d, x in zip(ranks, local)])
    assert np.array_equal(np.concatenate((rho, v)), ref)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, kind, halo, n, steps, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bc = "extend_both"
    road = _make_road(n, seed=5)
    d = DomainDecomposedRoad(n, rank, world, halo, _oracle_stepper(kind, bc))
    local = d.scatter(road)
    for _ in range(steps):
        local = d.step(local)
    full = d.gather(torch.from_numpy(np.ascontiguousarray(local)))
    if rank == 0:
        q.put(full.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind,halo", [("arz", 1), ("arz", 5), ("lwr", 3)])
def test_decomposed_road_two_ranks_gloo(kind, halo):
    n, steps, world = 200, 17, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker,
                         args=(r, world, port, kind, halo, n, steps, q))
             for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = _make_road(n, seed=5)
    fn = o.arz_step if kind == "arz" else o.lwr_step
    for _ in range(steps):
        ref = fn(ref, boundary_conditions="extend_both")
    assert np.array_equal(got, ref)


def test_advance_runs_whole_blocks_and_refreshes():
    """advance(n) == n x step(); a stepper offering ``block`` gets one call per
    run of steps between exchanges."""
    n, halo, steps = 120, 4, 11
    road = _make_road(n, seed=9)

    def exchange(send_l, send_r):
        return None, None

    base = _oracle_stepper("lwr", "extend_both")
    d1 = DomainDecomposedRoad(n, 0, 1, halo, base, exchange=exchange)
    ref = d1.scatter(road)
    for _ in range(steps):
        ref = d1.step(ref)

    blocks = []

    def stepper(x, bl, br):
        return base(x, bl, br)

    def block(x, k, bl, br):
        blocks.append(k)
        for _ in range(k):
            x = base(x, bl, br)
        return x

    stepper.block = block
    d2 = DomainDecomposedRoad(n, 0, 1, halo, stepper, exchange=exchange)
    out = d2.advance(d2.scatter(road), steps)
    assert np.array_equal(out, ref)
    assert blocks == [4, 4, 3]
    d2.advance(out, 2)                # finishes the open block first: 1 + 1
    assert blocks == [4, 4, 3, 1, 1]


def test_begin_starts_a_new_run_on_the_same_object():
    """begin() == scatter() plus a reset of the exchange state (without a
    mailbox there is no pending push to consume)."""
    n, halo = 120, 4
    road = _make_road(n, seed=11)
    base = _oracle_stepper("arz", "extend_both")
    d = DomainDecomposedRoad(n, 0, 1, halo, base, exchange=lambda a, b: (None, None))
    first = d.advance(d.begin(road), 6)
    assert d._since_exchange == 2
    again = d.advance(d.begin(road), 6)
    assert d._since_exchange == 2
    assert np.array_equal(first, again)
    assert np.array_equal(d.begin(road), d.scatter(road))
