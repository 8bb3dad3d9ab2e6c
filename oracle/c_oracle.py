"""ctypes front end of the C restatement (oracle/c_oracle.c).

TEST INFRASTRUCTURE (see oracle/__init__.py): a second checker and the CPU
baseline timed by bench.py.
"""
import ctypes
import os

import numpy as np

from oracle import build as _build

_BC = {"loop": 0, "extend": 1, "constant": 2, "halo": 3}


class _Params(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in
                ("dx", "dt", "rho_max", "v_max", "tau", "lam", "rho_crit")] + \
               [("bc_l", ctypes.c_int), ("bc_r", ctypes.c_int)] + \
               [(n, ctypes.c_double) for n in ("bl0", "bl1", "br0", "br1")]


_lib = None


def _load():
    global _lib
    if _lib is None:
        path = _build.OUTPUT if os.path.isfile(_build.OUTPUT) else _build.build()
        lib = ctypes.CDLL(path)
        lib.mt_oracle_step.restype = ctypes.c_int
        lib.mt_oracle_step.argtypes = [
            ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long,
            ctypes.c_long, ctypes.POINTER(_Params), ctypes.c_int]
        lib.mt_oracle_max_threads.restype = ctypes.c_int
        _lib = lib
    return _lib


def max_threads():
    return _load().mt_oracle_max_threads()


def _bc(model, bc):
    z = (0.0, 0.0)
    if bc == "loop":
        return 0, 0, z, z
    if bc == "extend_both":
        return 1, 1, z, z
    if isinstance(bc, dict) and list(bc)[0] == "constant_both":
        lo, hi = bc["constant_both"]
        if model == "arz":
            return 2, 2, tuple(lo), tuple(hi)
        return 2, 2, (lo, 0.0), (hi, 0.0)
    raise ValueError("Unknown boundary conditions: {}".format(bc))


def step(model, obs, dx=50, dt=0.5, rho_max=1, v_max=30, tau=9, lam=1,
         boundary_conditions="loop", rho_crit=None, out=None, n_threads=0):
    """One step of ``model`` ('lwr' | 'arz') on a (B, 2N) float64 batch."""
    obs = np.ascontiguousarray(obs, dtype=np.float64)
    squeeze = obs.ndim == 1
    batch = obs[None, :] if squeeze else obs
    b, n = batch.shape[0], batch.shape[1] // 2
    if out is None:
        out = np.empty_like(batch)
    cl, cr, lv, rv = _bc(model, boundary_conditions)
    p = _Params(dx=dx, dt=dt, rho_max=rho_max, v_max=v_max, tau=tau, lam=lam,
                rho_crit=(rho_max / 2 if rho_crit is None else rho_crit),
                bc_l=cl, bc_r=cr, bl0=lv[0], bl1=lv[1], br0=rv[0], br1=rv[1])
    rc = _load().mt_oracle_step(1 if model == "arz" else 0,
                                batch.ctypes.data, out.ctypes.data, b, n,
                                ctypes.byref(p), int(n_threads))
    if rc != 0:
        raise RuntimeError("mt_oracle_step failed: %d" % rc)
    return out[0] if squeeze else out
