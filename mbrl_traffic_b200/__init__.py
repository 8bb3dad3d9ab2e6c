"""mbrl_traffic_b200 -- B200-native engine for the batched finite-volume step
of mbrl-traffic's macroscopic traffic models.

Host-side mirror of ``mbrl_traffic/models`` (same class names, constructor
arguments and ``get_next_obs`` contract) over ``libmtraffic.so``
(include/mtraffic.h).  There is no CPU fallback.
"""
from mbrl_traffic_b200.base import Model
from mbrl_traffic_b200.lwr import LWRModel
from mbrl_traffic_b200.arz import ARZModel
from mbrl_traffic_b200.nonlocal_ import NonLocalModel
from mbrl_traffic_b200.params import LWR_MODEL_PARAMS
from mbrl_traffic_b200.params import ARZ_MODEL_PARAMS
from mbrl_traffic_b200.params import NONLOCAL_MODEL_PARAMS
from mbrl_traffic_b200.engine import Engine, PinnedArray
from mbrl_traffic_b200.optimizers import (NelderMead, CrossEntropyMethod,
                                          GeneticAlgorithm)

__all__ = [
    "Model", "LWRModel", "ARZModel", "NonLocalModel", "Engine",
    "PinnedArray", "NelderMead", "CrossEntropyMethod", "GeneticAlgorithm",
    "LWR_MODEL_PARAMS", "ARZ_MODEL_PARAMS", "NONLOCAL_MODEL_PARAMS",
]
