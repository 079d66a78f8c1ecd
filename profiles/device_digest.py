"""Digest of the batched kernel's results on fixed seeded batches: per workload the nit array and
sha256 of the x / fun bytes.  Run on a GPU box before and after a kernel change; equal digests =
bit-identical results.

    python profiles/device_digest.py gpurun_out/digest.json
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers  # noqa: E402


def l1(n, m):
    return dict(l1_ratios=(np.arange(m) + 1) / n, l1_shifts=np.arange(m))


WORKLOADS = {
    "FDS_n100_l1_fista": ("FDS", dict(n_features=100, **l1(100, 3)), -2, 2, 1024, dict(nesterov=True)),
    "FDS_n100_l1_ista": ("FDS", dict(n_features=100, **l1(100, 3)), -2, 2, 64, dict(nesterov=False)),
    "FDS_n10_box_fista": ("FDS", dict(n_features=10, bounds=(0, np.inf)), 0, 2, 256, dict(nesterov=True)),
    "FDS_n20_fista": ("FDS", dict(n_features=20), -2, 2, 256, dict(nesterov=True)),
    "JOS1_n5_fista": ("JOS1", dict(n_features=5), -2, 4, 1000, dict(nesterov=True)),
    "JOS1_n50_l1_fista": ("JOS1", dict(n_features=50, **l1(50, 2)), -2, 4, 1024, dict(nesterov=True)),
    "JOS1_n50_l1_newton": ("JOS1", dict(n_features=50, **l1(50, 2)), -2, 4, 256,
                           dict(nesterov=True, dual_solver="newton")),
    "TRIDIA_l1_fista": ("TRIDIA", l1(3, 3), -1, 1, 256, dict(nesterov=True)),
    "LFR1_n30_fista": ("LinearFunctionRank1", dict(n_features=30), -1, 1, 256, dict(nesterov=True)),
    "ZDT1_n50_ista": ("ZDT1", dict(n_features=50), 0, 0.01, 128, dict(nesterov=False)),
    "SD_fista": ("SD", dict(), 1.5, 3, 128, dict(nesterov=True)),
    "TOI4_l1_fista": ("TOI4", l1(4, 2), -2, 5, 128, dict(nesterov=True)),
}


def digest():
    out = {}
    for name, (cls, kw, lo, hi, ns, o) in WORKLOADS.items():
        prob = helpers.device_problem(cls, kw)
        rng = np.random.RandomState(1000)
        X0 = rng.uniform(lo, hi, size=(ns, prob.n_features))
        br = prob.minimize_proximal_gradient_batched(X0, tol_internal=1e-11, max_iter=200000, **o)
        out[name] = {
            "nit_sum": int(br.nit.sum()), "nit_max": int(br.nit.max()),
            "converged": int((br.status == 1).sum()),
            "nit_sha": hashlib.sha256(br.nit.tobytes()).hexdigest()[:16],
            "x_sha": hashlib.sha256(br.x.tobytes()).hexdigest()[:16],
            "fun_sha": hashlib.sha256(br.fun.tobytes()).hexdigest()[:16],
            "n_dual_sum": int(br.n_dual.sum()),
        }
    return out


if __name__ == "__main__":
    d = digest()
    with open(sys.argv[1], "w") as fh:
        json.dump(d, fh, indent=1, sort_keys=True)
    print(json.dumps(d, sort_keys=True))
