"""Shared pytest configuration: the ``gpu`` marker and common fixtures."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_step.npz")


def pytest_configure(config):
    config.addinivalue_line(
        "markers", "gpu: needs a CUDA device (run on the B200 box only)")


class GoldenCase(object):
    """One frozen input/output pair produced by the reference itself."""

    def __init__(self, key, data):
        self.key = key
        self.obs = data[key + "/obs"]
        self.out1 = data[key + "/out1"]
        self.outn = data[key + "/outn"]
        self.meta = json.loads(str(data[key + "/meta"]))
        self.model = self.meta["model"]
        self.n_chain = self.meta["n_chain"]
        p = dict(self.meta["params"])
        bc = p["boundary_conditions"]
        if isinstance(bc, dict):  # JSON turned the tuples into lists
            v = bc["constant_both"]
            if isinstance(v[0], list):
                v = tuple(tuple(x) for x in v)
            else:
                v = tuple(v)
            p["boundary_conditions"] = {"constant_both": v}
        self.params = p

    def step_kwargs(self):
        """Keyword arguments for ``oracle.np_oracle.{lwr,arz}_step``."""
        keep = ["dx", "dt", "rho_max", "v_max", "lam", "boundary_conditions"]
        if self.model == "arz":
            keep.append("tau")
        return {k: self.params[k] for k in keep}


def load_golden():
    data = np.load(GOLDEN)
    keys = sorted({k.split("/")[0] for k in data.files})
    return [GoldenCase(k, data) for k in keys]


@pytest.fixture(scope="session")
def golden_cases():
    return load_golden()
