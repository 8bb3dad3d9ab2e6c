"""Worker of test_two_processes_exchange_through_cuda_ipc_on_one_gpu: one of
two ranks that share GPU 0 (not a test module: no ``test_`` prefix)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mbrl_traffic_b200 import _lib  # noqa: E402
from mbrl_traffic_b200.dist import DomainDecomposedRoad, PeerHalo, engine_stepper  # noqa: E402
from mbrl_traffic_b200.engine import Engine  # noqa: E402

PARAMS = dict(dx=50, dt=0.5, rho_max=1, v_max=30, lam=1, tau=9)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    torch.cuda.set_device(0)
    n, halo, steps = 12000, 8, 40
    rng = np.random.default_rng(77)
    rho = rng.uniform(0.05, 0.9, n)
    road_np = np.concatenate((rho, 30 * (1 - rho) + rng.normal(0, 1, n)))
    road = torch.as_tensor(road_np, device="cuda:0")
    ok = True
    for fused in (False, True):
        d = DomainDecomposedRoad(n, rank, world, halo, None)
        eng = Engine("arz", "float64", 1, d.n_local, device=0)
        d.stepper = engine_stepper(eng, _lib.MT_BC_EXTEND, _lib.MT_BC_EXTEND, **PARAMS)
        peer = PeerHalo(halo, "float64", 0).connect_group(rank, world)   # CUDA IPC handles
        d.peer = peer
        d.fuse_exchange = fused
        x = d.scatter(road).clone()
        x = d.advance(x, steps)
        torch.cuda.synchronize()
        timed_out = peer.timed_out()
        rho_own, v_own = d.own(x)
        mine = torch.cat((rho_own, v_own)).cpu()
        parts = [None] * world
        dist.all_gather_object(parts, mine.numpy())
        if rank == 0:
            full = Engine("arz", "float64", 1, n, device=0)
            full.set_params(boundary_conditions="extend_both", **PARAMS)
            a = road.reshape(1, -1).clone()
            ref = full.rollout(a, torch.empty_like(a), steps).cpu().numpy()[0]
            got = np.concatenate([p[:p.size // 2] for p in parts] +
                                 [p[p.size // 2:] for p in parts])
            same = bool(np.array_equal(got, ref))
            print("fused=%s identical=%s timed_out=%s" % (fused, same, timed_out), flush=True)
            ok = ok and same and not timed_out
        dist.barrier()
        peer.close()
    if rank == 0 and ok:
        print("IPC_OK", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
