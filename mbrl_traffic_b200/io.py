"""On-disk formats of the reference.

The reference's macroscopic CSV files are laid out ``time, speed_0..speed_{N-1},
density_0..density_{N-1}`` (``experiments/evaluate_model.py:76-102, 404-414``;
``tests/slow_tests/data/sample_macro.csv``) and its initial-condition files
``speed_* , density_*`` without the time column
(``experiments/pretrained/models/initial_conditions/*.csv``) -- **speed first**
-- while the models read ``[density | speed]`` (``lwr.py:197``,
``arz.py:210-211``).  Feeding the files as they are, as
``evaluate_model.py:383-401`` does, swaps the two halves; every reader here
returns observations in the models' ``[density | speed]`` order and the
writers put the reference's column order back.
"""
import csv

import numpy as np


def _columns(header):
    names = [h.strip().lstrip("#").strip() for h in header]
    speed = [i for i, h in enumerate(names) if h.startswith("speed_")]
    dens = [i for i, h in enumerate(names) if h.startswith("density_")]
    if len(speed) != len(dens) or not speed:
        raise ValueError("expected speed_k and density_k columns in equal number")
    speed.sort(key=lambda i: int(names[i].split("_")[1]))
    dens.sort(key=lambda i: int(names[i].split("_")[1]))
    time = names.index("time") if "time" in names else None
    return time, speed, dens


def read_macro_csv(path):
    """Read a reference macro file.  Returns ``(times, obs)`` with ``obs`` of
    shape ``(T, 2N)`` in the models' ``[density | speed]`` order; ``times`` is
    None for files without a time column (the initial-condition files)."""
    with open(path, newline="") as f:
        rows = list(csv.reader(f))
    t_col, speed, dens = _columns(rows[0])
    data = np.array([[float(x) for x in r] for r in rows[1:] if r], dtype=np.float64)
    obs = np.concatenate((data[:, dens], data[:, speed]), axis=1)
    times = data[:, t_col] if t_col is not None else None
    return times, obs


def write_macro_csv(path, times, obs):
    """Write ``obs`` (``(T, 2N)``, ``[density | speed]``) in the reference's
    column order ``time, speed_k..., density_k...``
    (``evaluate_model.py:404-414``)."""
    obs = np.asarray(obs, dtype=np.float64)
    n = obs.shape[1] // 2
    header = ["time"] + ["speed_%d" % k for k in range(n)] + \
        ["density_%d" % k for k in range(n)]
    table = np.concatenate((np.asarray(times, dtype=np.float64)[:, None],
                            obs[:, n:], obs[:, :n]), axis=1)
    np.savetxt(path, table, delimiter=",", header=",".join(header), comments="")


def load_initial_conditions(path, min_density=None):
    """Rows of an initial-condition file as a ``(B, 2N)`` batch of
    ``[density | speed]`` observations -- one road per row, ready for
    ``get_next_obs``.  ``min_density`` optionally lifts empty sections (the
    reference's snapshots contain exact zeros, outside the ARZ domain)."""
    _, obs = read_macro_csv(path)
    if min_density is not None:
        n = obs.shape[1] // 2
        obs[:, :n] = np.maximum(obs[:, :n], min_density)
    return obs


def rollout_to_csv(model, obs0, n_steps, path, dt=None):
    """Simulate one road for ``n_steps`` (the loop of
    ``evaluate_model.py:388-403``) and save the trajectory as a macro CSV."""
    # every frame stays on the device until the run is over: one upload, one
    # download (FiniteVolumeModel.trajectory), not a round trip per step
    frames = model.trajectory(np.asarray(obs0, dtype=np.float64), n_steps=n_steps)
    dt = model.dt if dt is None else dt
    times = np.arange(len(frames)) * dt
    write_macro_csv(path, times, frames)
    return times, frames
