"""Experimental: per-CTA timeline of the pipelined step kernel.

Needs a side build with -DMT_TRACE ("where the per-launch time goes"):
    nvcc ... -DMT_TRACE -o mbrl_traffic_b200/libmt_trace.so mbrl_traffic_b200/csrc/mtraffic.cu
    MT_LIB=$PWD/mbrl_traffic_b200/libmt_trace.so python tools/trace_pipe.py
Stamps (%globaltimer, ns) per CTA: 0 entry, 1 producer past the grid dependency,
2 first tile landed, 3 last load issued, 4 last tile landed, 5 last tile computed,
6 last store drained.  The last 8 launches are kept round-robin.
"""
import argparse
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
import json
import numpy as np
import torch
from mbrl_traffic_b200.engine import Engine
from mbrl_traffic_b200 import _lib


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--cells", type=int, default=1000)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--chained", action="store_true")
    ap.add_argument("--per-cta", action="store_true")
    args = ap.parse_args()
    lib = _lib.load()
    dev = torch.device("cuda:0")
    trace = torch.zeros(8 * 1024 * 8, dtype=torch.int64, device=dev)
    lib.mt_debug_trace.argtypes = [ctypes.c_void_p]
    eng = Engine("arz", "float64", args.envs, args.cells)
    eng.set_params(dx=50.0, dt=0.5, rho_max=1.0, v_max=30.0, tau=9.0, lam=1.0, rho_crit=0.5,
                   boundary_conditions="loop")
    g = torch.Generator(device=dev).manual_seed(0)
    n_sets = 4
    ins = [torch.cat([0.1 + 0.8 * torch.rand(args.envs, args.cells, generator=g, device=dev, dtype=torch.float64),
                      30 * torch.rand(args.envs, args.cells, generator=g, device=dev, dtype=torch.float64)], 1)
           for _ in range(n_sets)]
    outs = [torch.empty_like(x) for x in ins]
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for k in range(8):
            eng.step(ins[k % n_sets], outs[k % n_sets], stream=s.cuda_stream)
        s.synchronize()
        lib.mt_debug_trace(ctypes.c_void_p(trace.data_ptr()))
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            for k in range(args.steps):
                if args.chained:
                    a, b = (ins[0], outs[0]) if k % 2 == 0 else (outs[0], ins[0])
                else:
                    a, b = ins[k % n_sets], outs[k % n_sets]
                eng.step(a, b, stream=s.cuda_stream)
        graph.replay()
        s.synchronize()
    tr = trace.cpu().numpy().reshape(8, 1024, 8)
    # order the 8 kept launches by their start time
    starts = [tr[j][tr[j][:, 0] > 0][:, 0].min() for j in range(8)]
    order = np.argsort(starts)
    rows = []
    prev_end = None
    for j in order:
        t = tr[j]
        live = t[:, 0] > 0
        t = t[live].astype(np.int64)
        t0 = t[:, 0].min()
        rel = t - t0
        row = {"ctas": int(live.sum()),
               "gap_from_prev_end_ns": None if prev_end is None else int(t0 - prev_end)}
        names = ["entry", "dep_passed", "first_tile_landed", "last_load_issued",
                 "last_tile_landed", "last_tile_computed", "last_store_drained"]
        for k, nm in enumerate(names):
            col = rel[:, k]
            row[nm] = {"min": int(col.min()), "med": int(np.median(col)), "max": int(col.max())}
        row["span_ns"] = int(t[:, 6].max() - t0)
        prev_end = t[:, 6].max()
        rows.append(row)
    for r in rows[2:]:
        print(json.dumps(r))
    if args.per_cta:
        t = tr[order[-2]]
        live = t[:, 0] > 0
        t = t[live].astype(np.int64)
        t0 = t[:, 0].min()
        fin = t[:, 6] - t0
        smid = t[:, 7]
        idx = np.nonzero(live)[0]
        by = np.argsort(fin)
        print("finish order (block, smid, entry, finish ns):")
        for j in by:
            print(int(idx[j]), int(smid[j]), int(t[j, 0] - t0), int(fin[j]))


if __name__ == "__main__":
    main()
