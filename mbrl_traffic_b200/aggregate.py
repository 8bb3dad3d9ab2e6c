"""Microscopic -> macroscopic aggregation on the GPU.

Restates the reference's notebook loop
(``experiments/plotting_results.ipynb`` cell 8): vehicle records
``(time, global_position, speed)`` are binned into ``DT`` time rows and ``DX``
sections; per bin the mean speed and the density ``count / DX`` are returned,
empty sections get the free-flow speed.  The result is bit-equal to the
reference's sequential accumulation (``mt_aggregate``, no float atomics).
"""
import numpy as np

from mbrl_traffic_b200 import _lib


def aggregate_micro(time, position, speed, dx=50.0, dt=0.5, length=1500.0,
                    total_time=1800.0, empty_speed=30.0, device=0):
    """Returns ``(speeds, densities)``, device tensors of shape
    ``(round(total_time/dt) + 1, round(length/dx))`` (cell 8's array shapes).
    Inputs may be NumPy arrays or CUDA tensors of equal length."""
    import torch
    dev = torch.device("cuda", device)
    cols = [torch.as_tensor(np.asarray(c, dtype=np.float64) if not hasattr(c, "data_ptr")
                            else c, device=dev).to(torch.float64).contiguous()
            for c in (time, position, speed)]
    t, x, v = cols
    m = t.numel()
    n_rows = int(round(total_time / dt)) + 1
    n_sections = int(round(length / dx))
    # the kernel wants the records grouped by time row; emission files are,
    # anything else is stable-sorted so that the per-bin order is preserved
    rows = torch.floor(t / dt)
    rows = torch.where(rows < 0, rows + n_rows, rows)        # Python negative index
    rows = torch.where((rows < 0) | (rows >= n_rows) | torch.isnan(rows),
                       torch.full_like(rows, n_rows), rows)  # dropped: sentinel row
    if m > 1 and bool((rows[1:] < rows[:-1]).any()):
        order = torch.sort(rows, stable=True).indices
        t, x, v = t[order].contiguous(), x[order].contiguous(), v[order].contiguous()
    scratch = torch.empty(2 * max(m, 1) + 1, dtype=torch.int32, device=dev)
    speeds = torch.empty(n_rows, n_sections, dtype=torch.float64, device=dev)
    dens = torch.empty_like(speeds)
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(_lib.load().mt_aggregate(
        t.data_ptr(), x.data_ptr(), v.data_ptr(), m, float(dt), float(dx),
        n_rows, n_sections, float(empty_speed), scratch.data_ptr(),
        speeds.data_ptr(), dens.data_ptr(), device, stream))
    return speeds, dens


def macro_observations(speeds, densities):
    """Rows of ``[density | speed]`` observations, the layout the models read
    (the reference's files are speed-first; see ``io.py``)."""
    import torch
    return torch.cat((densities, speeds), dim=1)
