"""Multi-rank checks, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N \
        --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py

1. road-batch sharding: each rank steps its shard; the gathered result is
   bit-identical to rank 0 stepping the whole batch.
2. one long ARZ road, domain-decomposed with NCCL halo exchange: identical to
   the single-GPU run at every cell; reports cell-updates/s.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mbrl_traffic_b200 import _lib  # noqa: E402
from mbrl_traffic_b200.dist import (DomainDecomposedRoad, PeerHalo, engine_stepper,  # noqa: E402
                                    shard_range)
from mbrl_traffic_b200.engine import Engine  # noqa: E402

PARAMS = dict(dx=50, dt=0.5, rho_max=1, v_max=30, lam=1, tau=9)


def make_obs(b, n, seed):
    rng = np.random.default_rng(seed)
    rho = rng.uniform(0.05, 0.9, (b, n))
    v = 30 * (1 - rho) + rng.normal(0, 1, (b, n))
    return np.concatenate((rho, v), axis=1)


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    report = {"world": world}

    # ---- 1. road-batch sharding ------------------------------------------
    b, n, steps = 1024, 1000, 10
    obs = torch.as_tensor(make_obs(b, n, 1), device=dev)
    lo, hi = shard_range(b, rank, world)
    eng = Engine("arz", "float64", hi - lo, n, device=local_rank)
    eng.set_params(boundary_conditions="loop", **PARAMS)
    mine = eng.rollout(obs[lo:hi].clone(), torch.empty_like(obs[lo:hi]), steps)
    parts = [torch.empty_like(obs[slice(*shard_range(b, r, world))]) for r in range(world)]
    if len({p.shape[0] for p in parts}) == 1:
        dist.all_gather(parts, mine.contiguous())
    else:
        for r in range(world):
            if r == rank:
                parts[r].copy_(mine)
            dist.broadcast(parts[r], src=r)
    gathered = torch.cat(parts)
    whole = Engine("arz", "float64", b, n, device=local_rank)
    whole.set_params(boundary_conditions="loop", **PARAMS)
    ref = whole.rollout(obs.clone(), torch.empty_like(obs), steps)
    report["sharded_bit_identical"] = bool(torch.equal(gathered, ref))

    # ---- 2. one long road ------------------------------------------------
    n_road = int(os.environ.get("MT_ROAD_CELLS", 10_000_000))
    halo = int(os.environ.get("MT_HALO", 32))
    steps = 8 * halo
    road = torch.as_tensor(make_obs(1, n_road, 2)[0], device=dev)
    d = DomainDecomposedRoad(n_road, rank, world, halo, None)
    eng = Engine("arz", "float64", 1, d.n_local, device=local_rank)
    d.stepper = engine_stepper(eng, _lib.MT_BC_EXTEND, _lib.MT_BC_EXTEND, **PARAMS)
    peer = PeerHalo(halo, "float64", local_rank).connect_group(rank, world)

    def timed_run(use_peer):
        d.peer = peer if use_peer else None
        x = d.scatter(road).clone()
        d._since_exchange = 0
        x = d.advance(x, 2 * halo + 1)        # warm-up incl. two exchanges (NCCL set-up)
        x = d.scatter(road).clone()
        d._since_exchange = 0
        dist.barrier()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        x = d.advance(x, steps)               # one mt_rollout call per run of `halo` steps
        e1.record()
        torch.cuda.synchronize()
        tt = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return x, float(tt.item())

    local_nccl, t_nccl = timed_run(False)
    local, t_peer = timed_run(True)
    report["peer_halo_timed_out"] = bool(peer.timed_out())
    report["peer_equals_nccl"] = bool(torch.equal(local, local_nccl))
    full = d.gather(local)
    single = Engine("arz", "float64", 1, n_road, device=local_rank)
    single.set_params(boundary_conditions="extend_both", **PARAMS)
    x = road.reshape(1, -1).clone()
    t0 = time.perf_counter()
    ref = single.rollout(x, torch.empty_like(x), steps)
    torch.cuda.synchronize()
    t_single = time.perf_counter() - t0
    report["road_cells"] = n_road
    report["halo"] = halo
    report["road_identical_to_single_gpu"] = bool(torch.equal(full, ref.reshape(-1)))
    report["road_cell_updates_per_s"] = n_road * steps / t_peer
    report["road_cell_updates_per_s_nccl_halo"] = n_road * steps / t_nccl
    report["single_gpu_cell_updates_per_s"] = n_road * steps / t_single
    if rank == 0:
        print(json.dumps(report), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    ok = (report["sharded_bit_identical"] and report["road_identical_to_single_gpu"]
          and report["peer_equals_nccl"] and not report["peer_halo_timed_out"])
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
