"""Converged m >= 3 fixtures from the UNMODIFIED reference (/root/reference), at the options of
benchmarks/benchmark.py:303-372 (tol_internal = 1e-11, max_iter = 1e8, and the reference's
DEFAULT max_iter_internal = 100000 -- nothing capped).

    python tests/golden/make_golden_converged.py [--only NAME] [--jobs 6] [--overwrite]

Build container only (the GPU box has no /root/reference).  One file per case under
``tests/golden/converged/<problem>__<algo>.npz`` holding the starts and, PER START, what the
reference returned (x, fun, nit, success, status) plus its full allerrs / allfuns traces (ragged,
stored concatenated with offsets) and the wall time of the reference solve.

These complement tests/golden/<problem>__<algo>.npz (make_golden.py), which bound trust-constr at
40 outer / 1000 inner iterations: here trust-constr runs for as long as it takes, so the stored
``nit`` / ``fun`` are what a user of the reference actually gets on these problems.
"""
from __future__ import annotations

import argparse
import os
import sys
import time
import warnings
from concurrent.futures import ProcessPoolExecutor, as_completed

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "converged")
sys.path.insert(0, HERE)

AB_GRID = [
    (0.0, 0.0), (0.0, 1 / 8), (0.0, 1 / 4),
    (1 / 6, 1 / 144), (1 / 6, 37 / 288), (1 / 6, 1 / 4),
    (1 / 4, 1 / 64), (1 / 4, 17 / 128), (1 / 4, 1 / 4),
    (1 / 2, 1 / 16), (1 / 2, 5 / 32), (1 / 2, 1 / 4),
    (3 / 4, 9 / 64), (3 / 4, 25 / 128), (3 / 4, 1 / 4),
]


def _l1(n, m):
    # benchmarks/benchmark.py:439-440
    return dict(l1_ratios=(np.arange(m) + 1) / n, l1_shifts=np.arange(m))


# name -> (class, ctor kwargs, low, high, n_starts, algos); start ranges of benchmark.py:463-471
def cases():
    c = {}
    c["FDS_n5"] = ("FDS", dict(n_features=5), -2, 2, 6, ["ista", "fista"])
    c["FDS_n10"] = ("FDS", dict(n_features=10), -2, 2, 6, ["ista", "fista"])
    c["FDS_n10_l1"] = ("FDS", dict(n_features=10, **_l1(10, 3)), -2, 2, 6,
                       ["ista", "fista", "fista_ab0", "fista_ab7", "fista_ab13"])
    c["FDS_n10_box"] = ("FDS", dict(n_features=10, bounds=(0, np.inf)), 0, 2, 4, ["fista"])
    c["FDS_n20_l1"] = ("FDS", dict(n_features=20, **_l1(20, 3)), -2, 2, 4, ["fista"])
    c["TRIDIA"] = ("TRIDIA", dict(), -1, 1, 4, ["ista", "fista"])
    c["TRIDIA_l1"] = ("TRIDIA", _l1(3, 3), -1, 1, 4, ["ista", "fista"])
    c["LFR1_n30"] = ("LinearFunctionRank1", dict(n_features=30), -1, 1, 4, ["ista", "fista"])
    c["FDS_n50_l1"] = ("FDS", dict(n_features=50, **_l1(50, 3)), -2, 2, 2, ["fista"])
    c["FDS_n100_l1"] = ("FDS", dict(n_features=100, **_l1(100, 3)), -2, 2, 2, ["fista"])
    return c


def algo_options(algo):
    if algo == "ista":
        o = dict(nesterov=False)
    elif algo == "fista":
        o = dict(nesterov=True)
    elif algo.startswith("fista_ab"):
        o = dict(nesterov=True, nesterov_ratio=AB_GRID[int(algo[8:])])
    else:
        raise ValueError(algo)
    o.update(tol_internal=1e-11, max_iter=100000000)      # benchmark.py:310-311
    return o


def _solve(cls, kw, x0, opts):
    import refshim

    refshim.install()
    warnings.simplefilter("ignore")
    import zfista.problems as zp

    prob = getattr(zp, cls)(**kw)
    t0 = time.time()
    res = prob.minimize_proximal_gradient(x0, return_all=True, **opts)
    return dict(x=np.asarray(res.x, dtype=np.float64), fun=np.asarray(res.fun, dtype=np.float64),
                nit=int(res.nit), success=bool(res.success), status=int(res.status),
                allerrs=np.asarray(res.allerrs, dtype=np.float64).reshape(-1),
                allfuns=np.asarray(res.allfuns, dtype=np.float64), seconds=time.time() - t0)


def _save(path, cls, kw, X0, opts, outs):
    offs = np.concatenate([[0], np.cumsum([len(o["allerrs"]) for o in outs])])
    save = dict(
        problem=cls, x0=X0, x=np.stack([o["x"] for o in outs]),
        fun=np.stack([o["fun"] for o in outs]), nit=np.array([o["nit"] for o in outs]),
        success=np.array([o["success"] for o in outs]),
        status=np.array([o["status"] for o in outs]),
        ref_seconds=np.array([o["seconds"] for o in outs]),
        allerrs=np.concatenate([o["allerrs"] for o in outs]), trace_offsets=offs,
        # allfuns of start i: rows offs[i] + i .. offs[i+1] + i + 1 (one more row than allerrs)
        allfuns=np.concatenate([o["allfuns"].reshape(len(o["allerrs"]) + 1, -1) for o in outs]),
        opt_keys=np.array(sorted(opts)), opt_vals=np.array([repr(opts[k]) for k in sorted(opts)]))
    for k, v in kw.items():
        save["kw_" + k] = np.asarray(v, dtype=np.float64)
    np.savez(path, **save)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--jobs", type=int, default=6)
    ap.add_argument("--overwrite", action="store_true")
    a = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    pending = {}
    with ProcessPoolExecutor(max_workers=a.jobs) as ex:
        futs = {}
        for pname, (cls, kw, low, high, n, algos) in cases().items():
            rng = np.random.RandomState(1000 + sum(map(ord, pname)))
            nf = kw.get("n_features", 3 if cls == "TRIDIA" else None)
            X0 = rng.uniform(low, high, size=(n, nf))
            for algo in algos:
                case = f"{pname}__{algo}"
                path = os.path.join(OUT, case + ".npz")
                if (a.only and a.only not in case) or (os.path.exists(path) and not a.overwrite):
                    continue
                opts = algo_options(algo)
                pending[case] = dict(path=path, cls=cls, kw=kw, X0=X0, opts=opts, outs=[None] * n,
                                     left=n, t0=time.time())
                for i in range(n):
                    futs[ex.submit(_solve, cls, kw, X0[i], opts)] = (case, i)
        for fut in as_completed(futs):
            case, i = futs[fut]
            p = pending[case]
            p["outs"][i] = fut.result()
            p["left"] -= 1
            print(f"[converged] {case}[{i}]: nit={p['outs'][i]['nit']} ok={p['outs'][i]['success']} "
                  f"{p['outs'][i]['seconds']:.1f}s", flush=True)
            if p["left"] == 0:
                _save(p["path"], p["cls"], p["kw"], p["X0"], p["opts"], p["outs"])
                print(f"[converged] saved {case}: nit={[o['nit'] for o in p['outs']]}", flush=True)


if __name__ == "__main__":
    main()
