"""Generate tests/golden/*.npz by running the UNMODIFIED reference models.

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden

The reference's own tests hold no numeric vector for the step
(tests/fast_tests/test_models.py:74-98, 213-237 are ``pass  # TODO``), so the
golden vectors are outputs of the reference code itself, produced through
``oracle/ref_shim.py`` on seeded inputs.  Each case stores the input
observation, the reference output after 1 step and after ``n_chain`` chained
steps, and the parameters.  The files are what pins ``oracle/np_oracle.py``
(and through it the CUDA path) on machines without the reference.
"""
import json
import os
import sys
import warnings

import numpy as np

from oracle import ref_shim

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")

PARAM_SETS = [
    dict(rho_max=1, v_max=30, lam=1, tau=9),          # reference defaults
    dict(rho_max=0.2, v_max=30, lam=1, tau=9),        # realistic veh/m
    dict(rho_max=1, v_max=30, lam=2, tau=5),          # quadratic Greenshields
]
BCS = {
    "loop": ("loop", "loop"),
    "extend_both": ("extend_both", "extend_both"),
    # LWR takes scalar densities, ARZ takes (rho, y) pairs (arz.py:347-360)
    "constant_both": ({"constant_both": (0.3, 0.2)},
                      {"constant_both": ((0.3, 0.1), (0.2, -0.05))}),
}


def make_obs(n, rho_max, v_max, lam, seed):
    """Seeded input: rho ~ rho_max U(0.05, 0.9), v = V(rho) + N(0, 1)."""
    r = np.random.default_rng(seed)
    rho = rho_max * r.uniform(0.05, 0.9, n)
    v = v_max * (1 - rho / rho_max) ** lam + r.normal(0.0, 1.0, n)
    return np.concatenate((rho, v))


def _chain(model, obs, n_steps):
    cur = obs.copy()
    for _ in range(n_steps):
        cur = model.get_next_obs(cur.copy(), None)
    return cur


def main():
    if not ref_shim.available():
        sys.exit("reference checkout not present; nothing generated")
    warnings.filterwarnings("ignore")  # fsolve 'not making good progress'
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    cases = {}
    seed = 0

    for model_name, sizes in (("lwr", (8, 20, 30, 1000)),
                              ("arz", (8, 20, 30, 100))):
        for bc_name, (bc_lwr, bc_arz) in BCS.items():
            for n in sizes:
                for pi, p in enumerate(PARAM_SETS):
                    seed += 1
                    if model_name == "lwr":
                        bc = bc_lwr
                        kw = dict(dx=50, dt=0.5, rho_max=p["rho_max"],
                                  rho_max_max=1, v_max=p["v_max"],
                                  v_max_max=30, lam=p["lam"],
                                  boundary_conditions=bc)
                        model = ref_shim.reference_lwr(**kw)
                        n_chain = 1000
                    else:
                        bc = bc_arz
                        kw = dict(dx=50, dt=0.5, rho_max=p["rho_max"],
                                  rho_max_max=1, v_max=p["v_max"],
                                  v_max_max=30, tau=p["tau"], tau_max=25,
                                  lam=p["lam"], boundary_conditions=bc)
                        model = ref_shim.reference_arz(**kw)
                        n_chain = 1000 if n <= 30 else 100
                    obs = make_obs(n, p["rho_max"], p["v_max"], p["lam"], seed)
                    out1 = model.get_next_obs(obs.copy(), None)
                    outn = _chain(model, obs, n_chain)
                    key = "%s_%s_n%d_p%d" % (model_name, bc_name, n, pi)
                    cases[key] = dict(
                        obs=obs, out1=out1, outn=outn,
                        meta=json.dumps(dict(
                            model=model_name, n=n, n_chain=n_chain, seed=seed,
                            params={k: v for k, v in kw.items()})))
                    print(key, "finite:", bool(np.isfinite(outn).all()),
                          flush=True)

    # Realistic initial conditions: the 1500 m ring snapshots shipped with the
    # reference are [speed | density]; the models read [density | speed]
    # (SURVEY.md appendix A), so the halves are swapped on ingestion.
    ic_path = os.path.join(
        ref_shim.REFERENCE_ROOT, "experiments", "pretrained", "models",
        "initial_conditions", "60.csv")
    ics = np.genfromtxt(ic_path, delimiter=",", skip_header=1)[:8]
    half = ics.shape[1] // 2
    ics = np.concatenate((ics[:, half:], ics[:, :half]), axis=1)
    ics[:, :half] = np.maximum(ics[:, :half], 1e-3)  # ARZ domain: rho > 0
    for model_name in ("lwr", "arz"):
        kw = dict(dx=50, dt=0.5, rho_max=0.2, rho_max_max=1, v_max=30,
                  v_max_max=30, lam=1, boundary_conditions="loop")
        if model_name == "lwr":
            model = ref_shim.reference_lwr(**kw)
        else:
            kw.update(tau=9, tau_max=25)
            model = ref_shim.reference_arz(**kw)
        out1 = np.stack([model.get_next_obs(r.copy(), None) for r in ics])
        outn = np.stack([_chain(model, r, 200) for r in ics])
        key = "%s_ring_ic60" % model_name
        cases[key] = dict(
            obs=ics, out1=out1, outn=outn,
            meta=json.dumps(dict(model=model_name, n=half, n_chain=200,
                                 seed=-1, params=kw)))
        print(key, "finite:", bool(np.isfinite(outn).all()), flush=True)

    flat = {}
    for key, c in cases.items():
        for field in ("obs", "out1", "outn"):
            flat["%s/%s" % (key, field)] = c[field]
        flat["%s/meta" % key] = np.array(c["meta"])
    path = os.path.join(GOLDEN_DIR, "reference_step.npz")
    np.savez_compressed(path, **flat)
    print("wrote", path, "cases:", len(cases),
          "bytes:", os.path.getsize(path))


if __name__ == "__main__":
    main()
