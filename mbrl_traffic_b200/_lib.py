"""ctypes binding of libmtraffic.so (the C ABI in include/mtraffic.h).

The library is built in-tree by ``__graft_entry__.build()`` (or
``python -m mbrl_traffic_b200.build``).  There is no fallback: if the shared
object is missing or no CUDA device is present, the product path raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MT_LIB selects an alternative build of the same library (tuning experiments)
LIB_PATH = os.environ.get("MT_LIB") or os.path.join(_HERE, "libmtraffic.so")

MT_OK, MT_EINVAL, MT_ECUDA, MT_ENOMEM, MT_ESTATE = 0, -1, -2, -3, -4
MT_MODEL_LWR, MT_MODEL_ARZ, MT_MODEL_NONLOCAL = 0, 1, 2
MT_F32, MT_F64 = 0, 1
MT_BC_LOOP, MT_BC_EXTEND, MT_BC_CONSTANT, MT_BC_HALO = 0, 1, 2, 3
MT_SCHEME_GODUNOV, MT_SCHEME_LXF = 0, 1
MT_FLAG_FAST_MATH = 1
MT_FLAG_NO_PIPELINE = 2
MT_FLAG_INDEPENDENT_INPUT = 4
MT_FLAG_DEPENDENT_INPUT = 8
MT_FLAG_FSOLVE_ABORT = 16

# every symbol include/mtraffic.h declares
EXPORTS = (
    "mt_version", "mt_last_error", "mt_create", "mt_destroy", "mt_set_params",
    "mt_step", "mt_rollout", "mt_step_host", "mt_density_sqerr",
    "mt_launch_count", "mt_alloc_pinned", "mt_free_pinned", "mt_aggregate",
    "mt_nonfinite", "mt_halo_create", "mt_halo_destroy", "mt_halo_handle", "mt_halo_connect",
    "mt_halo_connect_local", "mt_halo_push", "mt_halo_pull", "mt_halo_status",
    "mt_halo_advance", "mt_step_obs16", "mt_actor_head",
)


class MtConfig(ctypes.Structure):
    _fields_ = [
        ("model", ctypes.c_int32), ("dtype", ctypes.c_int32),
        ("n_envs", ctypes.c_int64), ("n_cells", ctypes.c_int64),
        ("device", ctypes.c_int32), ("max_n_eta", ctypes.c_int32),
        ("host_staging", ctypes.c_int32), ("reserved", ctypes.c_int32),
    ]


class MtParams(ctypes.Structure):
    _fields_ = [
        ("dx", ctypes.c_double), ("dt", ctypes.c_double),
        ("rho_max", ctypes.c_double), ("v_max", ctypes.c_double),
        ("tau", ctypes.c_double), ("lam", ctypes.c_double),
        ("rho_crit", ctypes.c_double),
        ("rho_max_env", ctypes.c_void_p), ("v_max_env", ctypes.c_void_p),
        ("tau_env", ctypes.c_void_p), ("lam_env", ctypes.c_void_p),
        ("rho_crit_env", ctypes.c_void_p),
        ("bc_left", ctypes.c_int32), ("bc_right", ctypes.c_int32),
        ("bc_left_val", ctypes.c_double * 2),
        ("bc_right_val", ctypes.c_double * 2),
        ("gamma", ctypes.c_void_p),
        ("n_eta", ctypes.c_int32), ("flags", ctypes.c_int32),
        ("scheme", ctypes.c_int32), ("reserved", ctypes.c_int32),
        ("alpha", ctypes.c_double),
    ]


class MtError(RuntimeError):
    """A libmtraffic call returned a negative status."""

    def __init__(self, code, message):
        super(MtError, self).__init__("libmtraffic error %d: %s" % (code, message))
        self.code = code
        self.message = message


_lib = None


def load():
    """Load libmtraffic.so once and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            "%s not found: build it with `python -c 'import __graft_entry__ as "
            "g; g.build()'` (there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    lib.mt_version.restype = ctypes.c_int
    lib.mt_version.argtypes = []
    lib.mt_last_error.restype = ctypes.c_char_p
    lib.mt_last_error.argtypes = []
    lib.mt_create.restype = ctypes.c_int
    lib.mt_create.argtypes = [ctypes.POINTER(MtConfig), ctypes.POINTER(vp)]
    lib.mt_destroy.restype = None
    lib.mt_destroy.argtypes = [vp]
    lib.mt_set_params.restype = ctypes.c_int
    lib.mt_set_params.argtypes = [vp, ctypes.POINTER(MtParams), vp]
    lib.mt_step.restype = ctypes.c_int
    lib.mt_step.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.mt_step_obs16.restype = ctypes.c_int
    lib.mt_step_obs16.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.mt_actor_head.restype = ctypes.c_int
    lib.mt_actor_head.argtypes = [i32, i32, vp, vp, vp, vp, i64, i32, ctypes.c_double, vp, vp, vp]
    lib.mt_rollout.restype = ctypes.c_int
    lib.mt_rollout.argtypes = [vp, vp, vp, vp, vp, i32, vp,
                               ctypes.POINTER(i32)]
    lib.mt_step_host.restype = ctypes.c_int
    lib.mt_step_host.argtypes = [vp, vp, vp, vp, vp, i32]
    lib.mt_density_sqerr.restype = ctypes.c_int
    lib.mt_density_sqerr.argtypes = [vp, vp, vp, vp, vp]
    lib.mt_aggregate.restype = ctypes.c_int
    lib.mt_aggregate.argtypes = [vp, vp, vp, i64, ctypes.c_double, ctypes.c_double,
                                 i32, i32, ctypes.c_double, vp, vp, vp, i32, vp]
    lib.mt_launch_count.restype = ctypes.c_int
    lib.mt_launch_count.argtypes = [vp, ctypes.POINTER(i64)]
    lib.mt_alloc_pinned.restype = ctypes.c_int
    lib.mt_alloc_pinned.argtypes = [ctypes.c_uint64, ctypes.POINTER(vp)]
    lib.mt_free_pinned.restype = ctypes.c_int
    lib.mt_free_pinned.argtypes = [vp]
    lib.mt_nonfinite.restype = ctypes.c_int
    lib.mt_nonfinite.argtypes = [vp, vp, vp, vp]
    lib.mt_halo_create.restype = ctypes.c_int
    lib.mt_halo_create.argtypes = [i32, i32, i32, ctypes.POINTER(vp)]
    lib.mt_halo_destroy.restype = None
    lib.mt_halo_destroy.argtypes = [vp]
    lib.mt_halo_handle.restype = ctypes.c_int
    lib.mt_halo_handle.argtypes = [vp, ctypes.c_char_p]
    lib.mt_halo_connect.restype = ctypes.c_int
    lib.mt_halo_connect.argtypes = [vp, ctypes.c_char_p, ctypes.c_char_p]
    lib.mt_halo_connect_local.restype = ctypes.c_int
    lib.mt_halo_connect_local.argtypes = [vp, vp, vp]
    lib.mt_halo_push.restype = ctypes.c_int
    lib.mt_halo_push.argtypes = [vp, vp, i64, vp]
    lib.mt_halo_pull.restype = ctypes.c_int
    lib.mt_halo_pull.argtypes = [vp, vp, i64, vp]
    lib.mt_halo_advance.restype = ctypes.c_int
    lib.mt_halo_advance.argtypes = [vp, vp, vp, vp, i32, i32, vp, ctypes.POINTER(i32)]
    lib.mt_halo_status.restype = ctypes.c_int
    lib.mt_halo_status.argtypes = [vp, ctypes.POINTER(i32)]
    _lib = lib
    return lib


def check(code):
    """Raise MtError (or ValueError for bad arguments) on a failed call."""
    if code == MT_OK:
        return
    msg = load().mt_last_error().decode("utf-8", "replace")
    if code == MT_EINVAL:
        raise ValueError(msg)
    raise MtError(code, msg)
