"""Host-side contract tests that need no GPU: constructor attributes and
defaults, model.json schema, boundary-condition parsing, and that the C-ABI
library loads and exports every symbol include/mtraffic.h declares.

They mirror what the reference's tests pin
(tests/fast_tests/test_models.py:54-72, 100-162, 194-211, 239-298).
"""
import json
import os
import re

import numpy as np
import pytest

from mbrl_traffic_b200 import ARZModel, LWRModel, Model
from mbrl_traffic_b200 import ARZ_MODEL_PARAMS, LWR_MODEL_PARAMS
from mbrl_traffic_b200 import _lib
from mbrl_traffic_b200.engine import parse_boundary_conditions

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class DummyOptimizer(object):
    def __init__(self, **kw):
        self.kw = kw


def _params(base):
    p = dict(base)
    p.update(sess=1, ob_space=2, ac_space=3, replay_buffer=4, verbose=5,
             optimizer_cls=DummyOptimizer)
    return p


def test_model_base_init():
    m = Model(sess=1, ob_space=2, ac_space=3, replay_buffer=4, verbose=5)
    assert (m.sess, m.ob_space, m.ac_space, m.replay_buffer, m.verbose) == \
        (1, 2, 3, 4, 5)
    for name, args in (("initialize", ()), ("get_next_obs", (0, 0)),
                       ("update", ()), ("compute_loss", (0, 0, 0)),
                       ("get_td_map", ()), ("save", ("",)), ("load", ("",))):
        with pytest.raises(NotImplementedError):
            getattr(m, name)(*args)


def test_default_params_match_reference_values():
    assert LWR_MODEL_PARAMS == dict(
        dx=50, dt=0.5, rho_max=1, rho_max_max=1, v_max=30, v_max_max=30,
        stream_model="greenshield", lam=1, boundary_conditions="loop",
        optimizer_cls="GeneticAlgorithm")
    assert ARZ_MODEL_PARAMS == dict(
        dx=50, dt=0.5, rho_max=1, rho_max_max=1, v_max=30, v_max_max=30,
        tau=9, tau_max=25, lam=1, boundary_conditions="loop",
        optimizer_cls="GeneticAlgorithm")


def test_arz_init_attributes():
    m = ARZModel(**_params(ARZ_MODEL_PARAMS))
    for k in ("dx", "dt", "rho_max", "rho_max_max", "v_max", "v_max_max",
              "tau", "tau_max", "lam", "boundary_conditions"):
        assert getattr(m, k) == ARZ_MODEL_PARAMS[k]
    assert m.optimizer_cls is DummyOptimizer
    assert m.optimizer_params == {}
    assert m.rho_crit == 0.5
    assert m.optimizer.kw["param_low"] == [0, 0, 0, 0]
    assert m.optimizer.kw["param_high"] == [1, 30, 25, 4]
    assert m.get_td_map() == {}
    assert m.initialize() is None


def test_lwr_init_attributes():
    m = LWRModel(**_params(LWR_MODEL_PARAMS))
    for k in ("dx", "dt", "rho_max", "rho_max_max", "v_max", "v_max_max",
              "stream_model", "lam", "boundary_conditions"):
        assert getattr(m, k) == LWR_MODEL_PARAMS[k]
    assert m.optimizer.kw["param_low"] == [0, 0, 0]
    assert m.optimizer.kw["param_high"] == [1, 30, 4]


@pytest.mark.parametrize("cls,base,keys", [
    (ARZModel, ARZ_MODEL_PARAMS,
     ["dx", "dt", "rho_max", "rho_max_max", "v_max", "v_max_max", "tau",
      "tau_max", "lam", "boundary_conditions"]),
    (LWRModel, LWR_MODEL_PARAMS,
     ["dx", "dt", "rho_max", "rho_max_max", "v_max", "v_max_max",
      "stream_model", "lam", "boundary_conditions"]),
])
def test_save_and_load_schema(tmp_path, cls, base, keys):
    m = cls(**_params(base))
    for i, k in enumerate(keys):
        setattr(m, k, i + 1)
    m.save(str(tmp_path))
    with open(tmp_path / "model.json") as f:
        data = json.load(f)
    assert data == {k: i + 1 for i, k in enumerate(keys)}
    m2 = cls(**_params(base))
    m2.load(str(tmp_path / "model.json"))
    for i, k in enumerate(keys):
        assert getattr(m2, k) == i + 1
    if cls is ARZModel:
        assert m2.rho_crit == 0.5   # frozen at construction (arz.py:199)


def test_boundary_condition_parsing():
    L, E, C = _lib.MT_BC_LOOP, _lib.MT_BC_EXTEND, _lib.MT_BC_CONSTANT
    assert parse_boundary_conditions("loop", "lwr")[:2] == (L, L)
    assert parse_boundary_conditions("extend_both", "arz")[:2] == (E, E)
    assert parse_boundary_conditions({"constant_both": (0.3, 0.2)}, "lwr") == \
        (C, C, (0.3, 0.0), (0.2, 0.0))
    assert parse_boundary_conditions(
        {"constant_both": ((0.3, 0.1), (0.2, -0.05))}, "arz") == \
        (C, C, (0.3, 0.1), (0.2, -0.05))
    assert parse_boundary_conditions({"inflow_outflow": 0.25}, "lwr")[:3] == \
        (C, E, (0.25, 0.0))
    with pytest.raises(ValueError, match="Unknown boundary conditions: ring"):
        parse_boundary_conditions("ring", "lwr")
    with pytest.raises(ValueError, match="Unknown boundary condition: 7"):
        parse_boundary_conditions(7, "arz")


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    with open(os.path.join(ROOT, "include", "mtraffic.h")) as f:
        header = f.read()
    declared = set(re.findall(r"\b(mt_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mt_version() >= 100


def test_struct_sizes_match_header_layout():
    import ctypes
    assert ctypes.sizeof(_lib.MtConfig) == 40
    assert ctypes.sizeof(_lib.MtParams) == 7 * 8 + 5 * 8 + 8 + 32 + 8 + 8 + 8 + 8


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product path must fail loudly."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    m = LWRModel(**_params(LWR_MODEL_PARAMS))
    with pytest.raises(_lib.MtError) as exc:
        m.get_next_obs(np.full(40, 0.5), None)
    assert exc.value.code == _lib.MT_ECUDA
    assert "no CPU fallback" in str(exc.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mbrl_traffic_b200")
    for dirpath, _, files in os.walk(pkg):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, name)) as f:
                    text = f.read()
                assert "import oracle" not in text and \
                    "from oracle" not in text, name


def test_nonlocal_model_host_contract(tmp_path):
    from mbrl_traffic_b200 import NonLocalModel, NONLOCAL_MODEL_PARAMS
    from mbrl_traffic_b200.nonlocal_ import lookahead_weights
    from oracle import np_oracle
    m = NonLocalModel(**_params(NONLOCAL_MODEL_PARAMS))
    assert m.eta == 800 and m.kernel == "linear"
    g = m.weights()
    assert g.shape == (16,) and abs(g.sum() - 1) < 1e-14
    # the product's weights and the oracle's are computed independently
    assert np.array_equal(g, np_oracle.nonlocal_weights(800, 50, "linear"))
    assert np.array_equal(lookahead_weights(200, 50, "constant"),
                          np_oracle.nonlocal_weights(200, 50, "constant"))
    with pytest.raises(ValueError):
        lookahead_weights(120, 50)
    assert m.cfl_number() <= 1.0
    m.save(str(tmp_path))
    m2 = NonLocalModel(**_params(NONLOCAL_MODEL_PARAMS))
    m2.eta, m2.kernel = 100, "constant"
    m2.load(str(tmp_path / "model.json"))
    assert m2.eta == 800 and m2.kernel == "linear"


def test_header_is_plain_c_and_links_against_the_library(tmp_path):
    """include/mtraffic.h compiles as C99 (no C++ in the boundary) and a C
    program using it links against libmtraffic.so and reads its version."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    lib_dir = os.path.join(ROOT, "mbrl_traffic_b200")
    if not os.path.isfile(os.path.join(lib_dir, "libmtraffic.so")):
        pytest.skip("libmtraffic.so not built")
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include "mtraffic.h"\n'
                   'int main(void) { mt_config c; mt_params p; (void)c; (void)p;\n'
                   '  printf("%d %d %d\\n", mt_version(), (int)sizeof(mt_config), MT_IPC_HANDLE_BYTES);\n'
                   '  return 0; }\n')
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror",
                    "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", lib_dir, "-lmtraffic", "-Wl,-rpath," + lib_dir], check=True)
    out = subprocess.run([str(exe)], stdout=subprocess.PIPE, universal_newlines=True, check=True)
    version, cfg_size, handle = out.stdout.split()
    assert int(version) > 0 and int(cfg_size) == 40 and int(handle) == 64


def test_nonlocal_scheme_choice_and_schema(tmp_path):
    """scheme / alpha (round 2): validated at construction, kept by save / load,
    and the Lax-Friedrichs CFL number uses the viscosity."""
    from mbrl_traffic_b200 import NonLocalModel, NONLOCAL_MODEL_PARAMS
    with pytest.raises(ValueError, match="Unknown scheme"):
        NonLocalModel(scheme="upwind", **_params(NONLOCAL_MODEL_PARAMS))
    m = NonLocalModel(scheme="lxf", alpha=35.0, **_params(NONLOCAL_MODEL_PARAMS))
    assert m.scheme == "lxf" and m.alpha == 35.0
    g0 = float(m.weights()[0])
    assert abs(m.cfl_number() - 0.5 / 50 * (35.0 + g0 * 30)) < 1e-15
    assert m._engine_kwargs()["scheme"] == "lxf"
    m.save(str(tmp_path))
    m2 = NonLocalModel(**_params(NONLOCAL_MODEL_PARAMS))
    assert m2.scheme == "godunov" and m2.alpha is None
    m2.load(str(tmp_path / "model.json"))
    assert m2.scheme == "lxf" and m2.alpha == 35.0


def test_abi_flags_and_version():
    assert _lib.MT_FLAG_DEPENDENT_INPUT == 8 and _lib.MT_FLAG_FSOLVE_ABORT == 16
    assert (_lib.MT_SCHEME_GODUNOV, _lib.MT_SCHEME_LXF) == (0, 1)
    with open(os.path.join(ROOT, "include", "mtraffic.h")) as f:
        header = f.read()
    for name in ("MT_FLAG_DEPENDENT_INPUT = 8", "MT_FLAG_FSOLVE_ABORT = 16", "MT_SCHEME_LXF = 1",
                 "#define MT_VERSION 200"):
        assert name in header, name
    assert _lib.load().mt_version() == 200
