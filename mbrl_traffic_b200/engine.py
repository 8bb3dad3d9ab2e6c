"""Thin object wrapper over the C ABI (include/mtraffic.h).

``Engine`` owns one ``mt_engine`` handle.  Device buffers are passed as
``torch`` CUDA tensors (or raw integer addresses); host buffers as NumPy
arrays.  Nothing here computes: every method is one call into
``libmtraffic.so``.
"""
import ctypes

import numpy as np

from mbrl_traffic_b200 import _lib

_MODELS = {"lwr": _lib.MT_MODEL_LWR, "arz": _lib.MT_MODEL_ARZ,
           "nonlocal": _lib.MT_MODEL_NONLOCAL}
_DTYPES = {"float32": _lib.MT_F32, "float64": _lib.MT_F64}


def dtype_name(dtype):
    """Normalise a NumPy / torch / string dtype to 'float32' or 'float64'."""
    name = str(dtype).replace("torch.", "")
    if name in ("float32", "float64"):
        return name
    name = np.dtype(dtype).name
    if name not in _DTYPES:
        raise ValueError("unsupported dtype: {}".format(dtype))
    return name


def parse_boundary_conditions(bc, model, error_prefix=None):
    """Map the reference's ``boundary_conditions`` value to the ABI codes.

    Accepts the reference's three forms (lwr.py:237-264, arz.py:316-364):
    ``"loop"``, ``"extend_both"``, ``{"constant_both": (left, right)}`` where
    each side is a density (LWR, non-local) or a ``(density, relative flow)``
    pair (ARZ).  Extensions: ``{"inflow_outflow": left}`` (constant upstream,
    free outflow downstream) and ``{"codes": (left_code, right_code,
    left_val, right_val)}`` used by the domain-decomposed driver.

    Raises the reference's ``ValueError`` text for anything else.
    """
    if error_prefix is None:
        error_prefix = ("Unknown boundary condition: {}" if model == "arz"
                        else "Unknown boundary conditions: {}")
    zero = (0.0, 0.0)

    def pair(x):
        if model == "arz":
            return (float(x[0]), float(x[1]))
        return (float(x), 0.0)

    if isinstance(bc, str) and bc == "loop":
        return _lib.MT_BC_LOOP, _lib.MT_BC_LOOP, zero, zero
    if isinstance(bc, str) and bc == "extend_both":
        return _lib.MT_BC_EXTEND, _lib.MT_BC_EXTEND, zero, zero
    if isinstance(bc, dict) and len(bc) > 0:
        key = list(bc.keys())[0]
        if key == "constant_both":
            left, right = bc[key][0], bc[key][1]
            return (_lib.MT_BC_CONSTANT, _lib.MT_BC_CONSTANT,
                    pair(left), pair(right))
        if key == "inflow_outflow":
            return (_lib.MT_BC_CONSTANT, _lib.MT_BC_EXTEND,
                    pair(bc[key]), zero)
        if key == "codes":
            cl, cr, lv, rv = bc[key]
            return int(cl), int(cr), tuple(lv), tuple(rv)
    raise ValueError(error_prefix.format(bc))


def _ptr(x):
    """Address of a torch tensor / NumPy array / int, or None."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    raise TypeError("cannot take the address of {!r}".format(type(x)))


def _stream_ptr(stream):
    """cudaStream_t of a torch stream (None -> torch's current stream)."""
    if isinstance(stream, int):
        return stream
    if stream is None:
        import torch
        stream = torch.cuda.current_stream()
    return stream.cuda_stream


class Engine(object):
    """One batched model instance on one GPU (``mt_create`` .. ``mt_destroy``)."""

    def __init__(self, model, dtype, n_envs, n_cells, device=0, max_n_eta=0,
                 host_staging=False):
        self._lib = _lib.load()
        self.model = model
        self.dtype = dtype_name(dtype)
        self.n_envs = int(n_envs)
        self.n_cells = int(n_cells)
        self.device = int(device)
        cfg = _lib.MtConfig(
            model=_MODELS[model], dtype=_DTYPES[self.dtype],
            n_envs=self.n_envs, n_cells=self.n_cells, device=self.device,
            max_n_eta=int(max_n_eta), host_staging=int(bool(host_staging)),
            reserved=0)
        handle = ctypes.c_void_p()
        _lib.check(self._lib.mt_create(ctypes.byref(cfg), ctypes.byref(handle)))
        self._h = handle
        self._keep = {}  # per-env parameter tensors must outlive the table build

    def close(self):
        if getattr(self, "_h", None):
            self._lib.mt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ #
    def set_params(self, dx, dt, rho_max, v_max, lam, boundary_conditions,
                   tau=1.0, rho_crit=None, gamma=None, fast_math=False,
                   no_pipeline=False, independent_input=False, dependent_input=False,
                   scheme="godunov", alpha=None, fsolve_abort=False, stream=None):
        """``mt_set_params``.  rho_max / v_max / tau / lam / rho_crit may be
        Python scalars or device tensors of shape (n_envs,) in the engine
        dtype; ``gamma`` is a device tensor of look-ahead weights."""
        p = _lib.MtParams()
        p.dx, p.dt = float(dx), float(dt)
        keep = {}

        def put(name, value, default=0.0):
            if value is None:
                setattr(p, name, default)
                setattr(p, name + "_env", None)
            elif hasattr(value, "data_ptr"):
                if value.numel() != self.n_envs:
                    raise ValueError("%s must have n_envs entries" % name)
                if dtype_name(value.dtype) != self.dtype:
                    raise ValueError("%s must be %s" % (name, self.dtype))
                value = value.contiguous()
                keep[name] = value
                setattr(p, name, 0.0)
                setattr(p, name + "_env", value.data_ptr())
            else:
                setattr(p, name, float(value))
                setattr(p, name + "_env", None)

        put("rho_max", rho_max)
        put("v_max", v_max)
        put("tau", tau, 1.0)
        put("lam", lam)
        put("rho_crit", rho_crit, -1.0)
        cl, cr, lv, rv = parse_boundary_conditions(boundary_conditions,
                                                   self.model)
        p.bc_left, p.bc_right = cl, cr
        p.bc_left_val[0], p.bc_left_val[1] = lv
        p.bc_right_val[0], p.bc_right_val[1] = rv
        if gamma is not None:
            keep["gamma"] = gamma
            p.gamma = gamma.data_ptr()
            p.n_eta = int(gamma.numel())
        if scheme not in ("godunov", "lxf"):
            raise ValueError("Unknown scheme: {}".format(scheme))
        p.scheme = _lib.MT_SCHEME_LXF if scheme == "lxf" else _lib.MT_SCHEME_GODUNOV
        p.alpha = -1.0 if alpha is None else float(alpha)
        p.flags = (_lib.MT_FLAG_FAST_MATH if fast_math else 0) | \
            (_lib.MT_FLAG_NO_PIPELINE if no_pipeline else 0) | \
            (_lib.MT_FLAG_INDEPENDENT_INPUT if independent_input else 0) | \
            (_lib.MT_FLAG_DEPENDENT_INPUT if dependent_input else 0) | \
            (_lib.MT_FLAG_FSOLVE_ABORT if fsolve_abort else 0)
        _lib.check(self._lib.mt_set_params(self._h, ctypes.byref(p),
                                           _stream_ptr(stream)))
        self._keep = keep

    def step(self, obs_in, obs_out, action=None, reward=None, stream=None, obs16=None):
        """``mt_step`` on device buffers; with ``obs16`` (a bfloat16 device
        tensor shaped like ``obs_out``) ``mt_step_obs16``: the step also leaves
        a bfloat16 copy of the new state there."""
        if obs16 is not None:
            _lib.check(self._lib.mt_step_obs16(
                self._h, _ptr(obs_in), _ptr(action), _ptr(obs_out), _ptr(reward),
                _ptr(obs16), _stream_ptr(stream)))
            return
        _lib.check(self._lib.mt_step(self._h, _ptr(obs_in), _ptr(action),
                                     _ptr(obs_out), _ptr(reward),
                                     _stream_ptr(stream)))

    def rollout(self, obs_a, obs_b, n_steps, actions=None, rewards=None,
                stream=None):
        """``mt_rollout``; returns the buffer that holds the final state."""
        flag = ctypes.c_int32(0)
        _lib.check(self._lib.mt_rollout(
            self._h, _ptr(obs_a), _ptr(obs_b), _ptr(actions), _ptr(rewards),
            int(n_steps), _stream_ptr(stream), ctypes.byref(flag)))
        return obs_b if flag.value else obs_a

    def step_host(self, obs_in, obs_out, action=None, reward=None, n_steps=1):
        """``mt_step_host`` on NumPy (ideally pinned) buffers; synchronous."""
        _lib.check(self._lib.mt_step_host(
            self._h, _ptr(obs_in), _ptr(action), _ptr(obs_out), _ptr(reward),
            int(n_steps)))

    def nonfinite(self, obs, flags_out, stream=None):
        """``mt_nonfinite``: ``flags_out`` (int32 ``(B,)`` device tensor) gets 1
        for every road whose state holds an inf or a NaN."""
        _lib.check(self._lib.mt_nonfinite(self._h, _ptr(obs), _ptr(flags_out),
                                          _stream_ptr(stream)))

    def density_sqerr(self, obs_pred, obs_target, out, stream=None):
        """``mt_density_sqerr`` on device buffers."""
        _lib.check(self._lib.mt_density_sqerr(
            self._h, _ptr(obs_pred), _ptr(obs_target), _ptr(out),
            _stream_ptr(stream)))

    @property
    def launch_count(self):
        out = ctypes.c_int64(0)
        _lib.check(self._lib.mt_launch_count(self._h, ctypes.byref(out)))
        return out.value


def actor_head(hidden, weight, bias, action_out, scale, stream=None, reward_prev=None,
               return_sum=None):
    """``mt_actor_head``: ``action_out[b] = scale * (tanh(hidden[b] . weight +
    bias) + 1)`` -- the output layer of a speed-limit policy written straight
    into the engine-dtype buffer ``mt_step`` reads its actions from.  ``hidden``
    (B, H), ``weight`` (H,) / (1, H), ``bias`` (1,): contiguous bfloat16 device
    tensors; ``action_out``: (B,) float32 / float64 device tensor.  With
    ``reward_prev`` / ``return_sum`` (both (B,) in the dtype of ``action_out``)
    the launch also does ``return_sum += reward_prev``."""
    lib = _lib.load()
    b, h = hidden.shape
    if weight.numel() != h or bias.numel() != 1 or action_out.numel() != b:
        raise ValueError("actor_head: shapes do not match")
    _lib.check(lib.mt_actor_head(
        hidden.device.index or 0, _DTYPES[dtype_name(action_out.dtype)], _ptr(hidden),
        _ptr(weight), _ptr(bias), _ptr(action_out), int(b), int(h), float(scale),
        _ptr(reward_prev), _ptr(return_sum), _stream_ptr(stream)))
    return action_out


class PinnedArray(object):
    """A page-locked NumPy array obtained through ``mt_alloc_pinned``."""

    def __init__(self, shape, dtype):
        self._lib = _lib.load()
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        p = ctypes.c_void_p()
        _lib.check(self._lib.mt_alloc_pinned(max(nbytes, 1), ctypes.byref(p)))
        self._p = p
        buf = (ctypes.c_char * max(nbytes, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=dtype,
                                   count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self._p:
            self.array = None
            self._lib.mt_free_pinned(self._p)
            self._p = None


class _PinnedBlock(object):
    """Owner object of one page-locked block handed out by ``PinnedPool``; the
    NumPy array built on it (and every view of that array) keeps it alive
    through ``ndarray.base``.  When the last of them is gone the block goes
    back to the pool."""

    def __init__(self, pool, ptr, nbytes, shape, dtype):
        self._pool, self._ptr, self._nbytes = pool, ptr, nbytes
        self.__array_interface__ = {
            "shape": tuple(shape), "typestr": np.dtype(dtype).str,
            "data": (ptr, False), "version": 3}

    def __del__(self):
        pool = self._pool
        if pool is not None:
            pool._give_back(self._ptr, self._nbytes)


class PinnedPool(object):
    """Page-locked result buffers for the host path of ``get_next_obs``.

    ``cudaMallocHost`` costs milliseconds, a device-to-host copy into pageable
    memory runs at a fraction of the PCIe rate; the reference contract wants a
    NEW array from every call (base.py:58-74).  ``take`` returns a fresh NumPy
    array on a page-locked block; the block is recycled once the caller has
    dropped the array and all of its views, so a loop that keeps only the
    latest observation allocates twice and then never again, and an array the
    caller still holds is never overwritten.  At most ``keep`` idle blocks per
    size are kept; the rest are freed.
    """

    def __init__(self, keep=3):
        self._lib = _lib.load()
        self._free = {}
        self._keep = keep

    def take(self, shape, dtype):
        dtype = np.dtype(dtype)
        nbytes = max(int(np.prod(shape)) * dtype.itemsize, 1)
        idle = self._free.get(nbytes)
        if idle:
            ptr = idle.pop()
        else:
            p = ctypes.c_void_p()
            _lib.check(self._lib.mt_alloc_pinned(nbytes, ctypes.byref(p)))
            ptr = p.value
        return np.asarray(_PinnedBlock(self, ptr, nbytes, shape, dtype))

    def _give_back(self, ptr, nbytes):
        idle = self._free.setdefault(nbytes, [])
        if len(idle) < self._keep:
            idle.append(ptr)
        else:
            try:
                self._lib.mt_free_pinned(ctypes.c_void_p(ptr))
            except Exception:
                pass

    def drain(self):
        """Free every idle block."""
        for idle in self._free.values():
            while idle:
                self._lib.mt_free_pinned(ctypes.c_void_p(idle.pop()))
