"""Differential dump for A/B runs of two library builds (ZFISTA_B200_LIB): device functors,
single subproblems (Brent and Newton) and traced batched solves on small seeded inputs.
python profiles/diff_dump.py out.npz ; python profiles/diff_dump.py --compare a.npz b.npz"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from device_digest import ROOT  # noqa: E402,F401
import helpers  # noqa: E402


def l1(n, m):
    return dict(l1_ratios=(np.arange(m) + 1) / n, l1_shifts=np.arange(m))


CASES = {
    "JOS1_n50_l1": ("JOS1", dict(n_features=50, **l1(50, 2)), -2, 4),
    "JOS1_n100_l1": ("JOS1", dict(n_features=100, **l1(100, 2)), -2, 4),
    "FDS_n10": ("FDS", dict(n_features=10), -2, 2),
    "FDS_n10_box": ("FDS", dict(n_features=10, bounds=(0, np.inf)), 0, 2),
    "FDS_n100_l1": ("FDS", dict(n_features=100, **l1(100, 3)), -2, 2),
}

if sys.argv[1] == "--compare":
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    for k in sorted(a.files):
        if a[k].tobytes() == b[k].tobytes():
            continue
        av, bv = a[k], b[k]
        neq = np.argwhere(~((av == bv) | (np.isnan(av.astype(float)) & np.isnan(bv.astype(float)))))
        first = tuple(neq[0]) if len(neq) else None
        print(f"{k}: {len(neq)} of {av.size} differ; first {first}: {av[first] if first else ''!r} vs "
              f"{bv[first] if first else ''!r}")
    print("compare done")
    sys.exit(0)

from zfista_b200 import solve_subproblems  # noqa: E402

out = {}
for name, (cls, kw, lo, hi) in CASES.items():
    prob = helpers.device_problem(cls, kw)
    n, m = prob.n_features, prob.n_objectives
    rng = np.random.RandomState(42)
    X = rng.uniform(lo, hi, size=(8, n))
    out[f"{name}.f"] = np.stack([prob.f(x) for x in X])
    out[f"{name}.jac"] = np.stack([prob.jac_f(x) for x in X])
    XO = X + 0.05 * rng.standard_normal(X.shape)
    if "bounds" in kw:
        XO = np.clip(XO, 0, None)
    LR = np.array([1.0, 0.5, 0.25, 1 / 64, 1 / 1024, 1e-5, 1e-6, 1e-7])
    for ds in ("reference", "newton"):
        x, fun, w = solve_subproblems(prob, X, XO, LR, tol_internal=1e-11, dual_solver=ds)
        out[f"{name}.sub_{ds}.x"], out[f"{name}.sub_{ds}.fun"], out[f"{name}.sub_{ds}.w"] = x, fun, w
        br = prob.minimize_proximal_gradient_batched(X, nesterov=True, tol_internal=1e-11,
                                                     max_iter=50, return_all=True, dual_solver=ds,
                                                     trace_capacity=50)
        out[f"{name}.solve_{ds}.allfuns"] = br.allfuns
        out[f"{name}.solve_{ds}.allerrs"] = br.allerrs
        out[f"{name}.solve_{ds}.n_dual"] = br.n_dual
        out[f"{name}.solve_{ds}.x"] = br.x
np.savez(sys.argv[1], **out)
