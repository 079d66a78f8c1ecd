"""Full result arrays of the digest workloads -> npz (library chosen by ZFISTA_B200_LIB).
python profiles/dump_results.py out.npz ; python profiles/dump_results.py --compare a.npz b.npz"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from device_digest import WORKLOADS  # noqa: E402
import helpers  # noqa: E402

if sys.argv[1] == "--compare":
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    for k in sorted(a.files):
        if a[k].shape != b[k].shape:
            print(k, "shape differs")
            continue
        same = a[k].tobytes() == b[k].tobytes()
        if not same:
            d = np.abs(a[k].astype(np.float64) - b[k].astype(np.float64))
            idx = np.argwhere(d > 0)
            neq = np.argwhere(a[k].view(np.uint64 if a[k].dtype == np.float64 else a[k].dtype)
                              != b[k].view(np.uint64 if b[k].dtype == np.float64 else b[k].dtype))
            print(f"{k}: DIFFERENT max|d|={d.max():.3e} n_value_diff={len(idx)} n_bit_diff={len(neq)} "
                  f"first={neq[0].tolist() if len(neq) else None} "
                  f"a={a[k][tuple(neq[0])] if len(neq) else ''} b={b[k][tuple(neq[0])] if len(neq) else ''}")
    print("compare done")
    sys.exit(0)

out = {}
for name, (cls, kw, lo, hi, ns, o) in WORKLOADS.items():
    prob = helpers.device_problem(cls, kw)
    X0 = np.random.RandomState(1000).uniform(lo, hi, size=(ns, prob.n_features))
    br = prob.minimize_proximal_gradient_batched(X0, tol_internal=1e-11, max_iter=200000, **o)
    for f in ("x", "fun", "nit", "status", "lr", "nfev", "n_dual", "err"):
        out[f"{name}.{f}"] = getattr(br, f)
np.savez(sys.argv[1], **out)
