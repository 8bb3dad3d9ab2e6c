"""Golden vectors for roads that hold an exact-zero density: outputs of the
UNMODIFIED reference ARZ model (build container only, needs /root/reference):

    python -m oracle.make_golden_zero

With rho == 0 somewhere, ``P = Q * (y / rho)`` is NaN there (arz.py:411), so is
``rhs`` (arz.py:460), and ``scipy.optimize.fsolve`` (arz.py:462-463) hands back
its initial guess for the whole road: y is frozen for that step while the
density still advances.  ``np_oracle.arz_step(..., fsolve_abort=True)`` and the
engine's ``MT_FLAG_FSOLVE_ABORT`` mimic that; these vectors pin them.
One step per case (``out1``); inputs are seeded.
"""
import json
import os
import sys
import warnings

import numpy as np

from oracle import ref_shim

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")

BCS = {"loop": "loop", "extend_both": "extend_both",
       "constant_both": {"constant_both": ((0.3, 0.1), (0.2, -0.05))},
       "constant_zero": {"constant_both": ((0.0, 0.0), (0.2, 0.1))}}
PARAM_SETS = [dict(rho_max=1, v_max=30, lam=1, tau=9),
              dict(rho_max=0.2, v_max=25, lam=2, tau=5)]
# (cells, positions of the exact zeros); the last rows have none (no abort)
ZEROS = [(20, [7]), (30, [0]), (30, [29]), (20, [3, 4]), (20, [1]), (20, [18]), (8, [2]),
         (20, [0, 19]), (100, [50]), (20, []), (100, [])]


def main():
    if not ref_shim.available():
        sys.exit("reference checkout not present; nothing generated")
    warnings.filterwarnings("ignore")
    flat, n_cases = {}, 0
    for bc_name, bc in BCS.items():
        for pi, p in enumerate(PARAM_SETS):
            for zi, (n, zpos) in enumerate(ZEROS):
                r = np.random.default_rng(9000 + 100 * pi + zi)
                rho = p["rho_max"] * r.uniform(0.05, 0.9, n)
                v = p["v_max"] * (1 - rho / p["rho_max"]) ** p["lam"] + r.normal(0, 1, n)
                rho[zpos] = 0.0
                obs = np.concatenate((rho, v))
                kw = dict(dx=50, dt=0.5, rho_max=p["rho_max"], rho_max_max=1, v_max=p["v_max"],
                          v_max_max=30, tau=p["tau"], tau_max=25, lam=p["lam"],
                          boundary_conditions=bc)
                model = ref_shim.reference_arz(**kw)
                out1 = model.get_next_obs(obs.copy(), None)
                key = "arz_zero_%s_p%d_z%d" % (bc_name, pi, zi)
                flat[key + "/obs"] = obs
                flat[key + "/out1"] = out1
                flat[key + "/meta"] = np.array(json.dumps(dict(
                    n=n, zeros=zpos, params=kw)))
                n_cases += 1
    path = os.path.join(GOLDEN_DIR, "reference_zero_density.npz")
    np.savez_compressed(path, **flat)
    print("wrote", path, "cases:", n_cases, "bytes:", os.path.getsize(path))


if __name__ == "__main__":
    main()
