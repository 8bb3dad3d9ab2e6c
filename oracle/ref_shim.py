"""Load the UNMODIFIED reference models from /root/reference (oracle only).

``import mbrl_traffic.models`` fails in this image because
``mbrl_traffic/models/__init__.py:3`` pulls in TensorFlow through ``fcnet``.
This shim registers empty stand-in packages and then executes the three
files on the hot path straight from where they lie:

    /root/reference/mbrl_traffic/models/base.py
    /root/reference/mbrl_traffic/models/lwr.py
    /root/reference/mbrl_traffic/models/arz.py

No reference source is copied into this repository.  The GPU box has no
``/root/reference``; ``available()`` is False there and only the committed
golden vectors (``tests/golden``) are used.
"""
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MT_REFERENCE_ROOT", "/root/reference")
_MODELS_DIR = os.path.join(REFERENCE_ROOT, "mbrl_traffic", "models")

_loaded = {}


def available():
    """Return True if the reference checkout is present on this machine."""
    return os.path.isfile(os.path.join(_MODELS_DIR, "arz.py"))


class _NullOptimizer(object):
    """Stand-in for ``optimizer_cls``; the models build one in __init__."""

    def __init__(self, **kwargs):
        self.kwargs = kwargs

    def solve(self, *args, **kwargs):
        raise RuntimeError("the oracle shim does not fit parameters")


def _load():
    if _loaded:
        return _loaded
    if not available():
        raise RuntimeError(
            "reference checkout not found under %s" % REFERENCE_ROOT)
    for pkg in ("mbrl_traffic", "mbrl_traffic.models"):
        if pkg not in sys.modules:
            mod = types.ModuleType(pkg)
            mod.__path__ = []
            sys.modules[pkg] = mod
    for name in ("base", "lwr", "arz"):
        dotted = "mbrl_traffic.models." + name
        spec = importlib.util.spec_from_file_location(
            dotted, os.path.join(_MODELS_DIR, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[dotted] = mod
        spec.loader.exec_module(mod)
        _loaded[name] = mod
    return _loaded


def reference_lwr(**params):
    """Build the reference ``LWRModel`` (lwr.py:14) with the given params."""
    mods = _load()
    kw = dict(sess=None, ob_space=None, ac_space=None, replay_buffer=None,
              verbose=0, stream_model="greenshield",
              optimizer_cls=_NullOptimizer)
    kw.update(params)
    return mods["lwr"].LWRModel(**kw)


def reference_arz(**params):
    """Build the reference ``ARZModel`` (arz.py:13) with the given params."""
    mods = _load()
    kw = dict(sess=None, ob_space=None, ac_space=None, replay_buffer=None,
              verbose=0, optimizer_cls=_NullOptimizer)
    kw.update(params)
    return mods["arz"].ARZModel(**kw)


def reference_model_base():
    """Return the reference ``Model`` base class (base.py:4)."""
    return _load()["base"].Model
