"""HBM bandwidth of a B200 as a function of the read : write mix.

    python tools/hbm_mix_bench.py [--mb 2048] [--reps 20]

The roofline denominator (MEASURED_PEAKS.json) is a COPY: one byte written per
byte read.  The non-local kernel writes two bytes per byte read ([rho | v] out,
rho in), so this measures what the memory system sustains for other mixes with
plain torch element-wise kernels on buffers far larger than L2: fill (0 : 1),
float32 -> float64 conversion (1 : 2), copy (1 : 1), float64 -> float32
(2 : 1) and a sum (1 : 0).  CUDA events, median of `--reps`.
"""
import argparse
import json

import torch


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=2048, help="size of the float64 buffer")
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    n = args.mb * (1 << 20) // 8
    x64 = torch.rand(n, dtype=torch.float64, device="cuda")
    y64 = torch.empty_like(x64)
    x32 = torch.rand(n, dtype=torch.float32, device="cuda")
    y32 = torch.empty_like(x32)
    cases = {
        "fill (0:1)": (lambda: y64.fill_(1.5), 0, 8 * n),
        "f32->f64 (1:2)": (lambda: y64.copy_(x32), 4 * n, 8 * n),
        "copy f64 (1:1)": (lambda: y64.copy_(x64), 8 * n, 8 * n),
        "f64->f32 (2:1)": (lambda: y32.copy_(x64), 8 * n, 4 * n),
        "sum (1:0)": (lambda: x64.sum(), 8 * n, 0),
    }
    out = {}
    for name, (fn, rd, wr) in cases.items():
        t = timed(fn, args.reps)
        out[name] = dict(ms=round(t * 1e3, 4), total_GBps=round((rd + wr) / t / 1e9, 1),
                         read_GBps=round(rd / t / 1e9, 1), write_GBps=round(wr / t / 1e9, 1))
    print(json.dumps(dict(buffer_mb=args.mb, reps=args.reps, cases=out), indent=1))


if __name__ == "__main__":
    main()
