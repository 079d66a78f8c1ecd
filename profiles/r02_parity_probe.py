"""Round-2 probe (GPU box): batched CUDA solves of the m >= 3 problems against the CPU statement
of the same algorithm (oracle DeviceModel), per start: nit, max|dx|, max rel dF, and how far the
CPU model moves under a 1-ulp perturbation of x0 (its own rounding envelope).

    python profiles/r02_parity_probe.py > gpurun_out/r02_parity_probe.txt
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers  # noqa: E402
from oracle import zfista_oracle as zo  # noqa: E402

warnings.simplefilter("ignore")


def l1(n, m):
    return dict(l1_ratios=(np.arange(m) + 1) / n, l1_shifts=np.arange(m))


CASES = [
    ("FDS", dict(n_features=100, **l1(100, 3)), -2, 2, 16, ["fista", "ista"]),
    ("FDS", dict(n_features=10), -2, 2, 8, ["fista", "ista"]),
    ("FDS", dict(n_features=10, **l1(10, 3)), -2, 2, 8, ["fista", "ista", "ab7", "ab13"]),
    ("FDS", dict(n_features=10, bounds=(0, np.inf)), 0, 2, 8, ["fista", "ista"]),
    ("TRIDIA", dict(), -1, 1, 8, ["fista", "ista"]),
    ("TRIDIA", l1(3, 3), -1, 1, 8, ["fista", "ista"]),
    ("LinearFunctionRank1", dict(n_features=30), -1, 1, 8, ["fista", "ista"]),
    ("JOS1", dict(n_features=50, **l1(50, 2)), -2, 4, 8, ["newton2"]),
]


def main():
    for cls, kw, lo, hi, ns, algos in CASES:
        prob = helpers.device_problem(cls, kw)
        spec = helpers.oracle_spec(cls, kw)
        rng = np.random.RandomState(2000 + sum(map(ord, cls)) + prob.n_features)
        X0 = rng.uniform(lo, hi, size=(ns, prob.n_features))
        for algo in algos:
            opts = dict(tol_internal=1e-11, max_iter=200000, nesterov=algo != "ista")
            extra = {}
            if algo.startswith("ab"):
                opts["nesterov_ratio"] = helpers.AB_GRID[int(algo[2:])]
            if algo == "newton2":
                extra["dual_solver"] = "newton"
            br = prob.minimize_proximal_gradient_batched(X0, **opts, **extra)
            print(f"== {cls} {kw.get('n_features', '')} l1={'l1_ratios' in kw} "
                  f"box={'bounds' in kw} {algo}", flush=True)
            for i in range(ns):
                mk = lambda: zo.DeviceModel(newton_for_two=(algo == "newton2"))
                r = zo.minimize_proximal_gradient(spec, X0[i], subproblem=mk(), **opts)
                env_n, env_x = 0, 0.0
                for s in range(2):
                    sg = np.sign(np.random.RandomState(s).uniform(-1, 1, X0.shape[1]))
                    rp = zo.minimize_proximal_gradient(spec, X0[i] * (1 + 1.1e-16 * sg),
                                                       subproblem=mk(), **opts)
                    env_n = max(env_n, abs(rp["nit"] - r["nit"]))
                    env_x = max(env_x, float(np.max(np.abs(rp["x"] - r["x"]))))
                dx = float(np.max(np.abs(br.x[i] - r["x"])))
                dF = float(np.max(np.abs(br.fun[i] - r["fun"]) / np.maximum(1, np.abs(r["fun"]))))
                print(f"  [{i}] nit gpu={int(br.nit[i])} cpu={r['nit']} st={int(br.status[i])}/"
                      f"{r['status']} dx={dx:.2e} dF={dF:.2e} | cpu self-envelope dnit={env_n} "
                      f"dx={env_x:.2e}", flush=True)


if __name__ == "__main__":
    main()
