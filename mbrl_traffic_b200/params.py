"""Default model parameters.

Values of ``LWR_MODEL_PARAMS`` / ``ARZ_MODEL_PARAMS`` follow the reference's
``mbrl_traffic/utils/train.py:38-66, 73-103`` (that module cannot be imported
without TensorFlow and Flow, so the defaults are restated here and pinned by
tests/test_models_host.py).  ``NONLOCAL_MODEL_PARAMS`` has no reference
counterpart (the reference never implemented the model).
"""

LWR_MODEL_PARAMS = dict(
    dx=50,
    dt=0.5,
    rho_max=1,
    rho_max_max=1,
    v_max=30,
    v_max_max=30,
    stream_model="greenshield",
    lam=1,
    boundary_conditions="loop",
    optimizer_cls="GeneticAlgorithm",
)

ARZ_MODEL_PARAMS = dict(
    dx=50,
    dt=0.5,
    rho_max=1,
    rho_max_max=1,
    v_max=30,
    v_max_max=30,
    tau=9,
    tau_max=25,
    lam=1,
    boundary_conditions="loop",
    optimizer_cls="GeneticAlgorithm",
)

NONLOCAL_MODEL_PARAMS = dict(
    dx=50,
    dt=0.5,
    rho_max=1,
    rho_max_max=1,
    v_max=30,
    v_max_max=30,
    stream_model="greenshield",
    lam=1,
    # look-ahead distance in metres (a whole number of cells) and the shape
    # of the non-increasing weight w_eta on [0, eta]
    eta=800,
    kernel="linear",
    boundary_conditions="loop",
    optimizer_cls="GeneticAlgorithm",
)

# maximum value for the lambda term in the Greenshield model (lwr.py:11)
LAM_MAX = 4
