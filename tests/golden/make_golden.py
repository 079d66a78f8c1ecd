"""Generate golden fixtures by running the UNMODIFIED reference (/root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py [--only NAME] [--jobs 8]

Writes one ``tests/golden/<case>.npz`` per case.  Every case stores the inputs
(``x0`` etc.) and what the reference returned (``x``, ``fun``, ``nit``, ``success``)
so tests can replay the identical inputs through the oracle (CPU) and through the
CUDA path (GPU) without the reference being present.

The reference's jax/jaxopt imports are satisfied by tests/golden/refshim.py (numpy
restatements of jaxopt.prox.prox_lasso / jaxopt.projection.projection_box; see
that file).  Nothing else of the reference is altered.

Case naming:   <Problem>_<variant>__<algo>
  algo: ista (nesterov=False) | fista (nesterov=True, a,b=(0,1/4)) |
        fista_dep (deprecated=True) | fista_ab<i> (i-th momentum pair of the
        PGM_experiment_with_various_a_b notebook grid)
"""
from __future__ import annotations

import argparse
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refshim  # noqa: E402

refshim.install()

from joblib import Parallel, delayed  # noqa: E402

# (a, b) grid of examples/PGM_experiment_with_various_a_b.ipynb / cameraman.ipynb
AB_GRID = [
    (0.0, 0.0), (0.0, 1 / 8), (0.0, 1 / 4),
    (1 / 6, 1 / 144), (1 / 6, 37 / 288), (1 / 6, 1 / 4),
    (1 / 4, 1 / 64), (1 / 4, 17 / 128), (1 / 4, 1 / 4),
    (1 / 2, 1 / 16), (1 / 2, 5 / 32), (1 / 2, 1 / 4),
    (3 / 4, 9 / 64), (3 / 4, 25 / 128), (3 / 4, 1 / 4),
]


def _l1(n_features, n_objectives):
    # benchmarks/benchmark.py:439-440
    return (np.arange(n_objectives) + 1) / n_features, np.arange(n_objectives)


def problem_specs():
    """name -> (class name, ctor kwargs, low, high, n_starts).  Start ranges follow
    benchmarks/benchmark.py:463-471."""
    s2 = np.sqrt(2)
    specs = {}

    def add(name, cls, kw, low, high, n):
        specs[name] = (cls, kw, low, high, n)

    add("JOS1_n5", "JOS1", dict(n_features=5), -2, 4, 16)
    add("JOS1_n50", "JOS1", dict(n_features=50), -2, 4, 8)
    add("JOS1_n200", "JOS1", dict(n_features=200), -2, 4, 4)
    r, s = _l1(5, 2)
    add("JOS1_n5_l1", "JOS1", dict(n_features=5, l1_ratios=r, l1_shifts=s), -2, 4, 16)
    r, s = _l1(50, 2)
    add("JOS1_n50_l1", "JOS1", dict(n_features=50, l1_ratios=r, l1_shifts=s), -2, 4, 8)
    # notebook variant: l1_ratios=(1/n, 1/n/2), shifts (0, 1)
    add("JOS1_n50_l1nb", "JOS1",
        dict(n_features=50, l1_ratios=(1 / 50, 1 / 100), l1_shifts=(0, 1)), -2, 4, 8)
    add("SD", "SD", dict(), [1, s2, s2, 1], [3, 3, 3, 3], 16)
    add("ZDT1_n50", "ZDT1", dict(n_features=50), 0, 0.01, 8)
    add("TOI4", "TOI4", dict(), -2, 5, 16)
    r, s = _l1(4, 2)
    add("TOI4_l1", "TOI4", dict(l1_ratios=r, l1_shifts=s), -2, 5, 16)
    add("TRIDIA", "TRIDIA", dict(), -1, 1, 4)
    r, s = _l1(3, 3)
    add("TRIDIA_l1", "TRIDIA", dict(l1_ratios=r, l1_shifts=s), -1, 1, 4)
    add("LFR1_n30", "LinearFunctionRank1", dict(n_features=30), -1, 1, 4)
    add("FDS_n5", "FDS", dict(n_features=5), -2, 2, 4)
    add("FDS_n10", "FDS", dict(n_features=10), -2, 2, 4)
    r, s = _l1(10, 3)
    add("FDS_n10_l1", "FDS", dict(n_features=10, l1_ratios=r, l1_shifts=s), -2, 2, 4)
    add("FDS_n10_box", "FDS", dict(n_features=10, bounds=(0, np.inf)), 0, 2, 4)
    r, s = _l1(100, 3)
    add("FDS_n100_l1", "FDS", dict(n_features=100, l1_ratios=r, l1_shifts=s), -2, 2, 2)
    return specs


ALGOS = {
    "ista": dict(nesterov=False),
    "fista": dict(nesterov=True),
    "fista_dep": dict(nesterov=True, deprecated=True),
}
for _i, _ab in enumerate(AB_GRID):
    ALGOS[f"fista_ab{_i}"] = dict(nesterov=True, nesterov_ratio=_ab)

# which algos per problem (the a,b sweep only where cheap)
DEFAULT_ALGOS = ["ista", "fista", "fista_dep"]
SWEEP_PROBLEMS = {"JOS1_n5": [0, 4, 12, 14], "JOS1_n50": [0, 7, 13], "SD": [3, 9]}
HEAVY = {"FDS_n100_l1": ["fista"], "FDS_n10_l1": ["ista", "fista"],
         "FDS_n10_box": ["fista"], "LFR1_n30": ["ista", "fista"],
         "JOS1_n200": ["ista", "fista"], "TRIDIA": ["ista"], "TRIDIA_l1": ["ista"],
         "FDS_n5": ["ista", "fista"], "FDS_n10": ["ista", "fista"]}


def _make_problem(cls, kw):
    import zfista.problems as zp

    return getattr(zp, cls)(**kw)


def _solve_one(cls, kw, x0, opts, return_all, here=HERE):
    # joblib workers are fresh processes: install the jax/jaxopt shim there too
    if here not in sys.path:
        sys.path.insert(0, here)
    import refshim as _shim

    _shim.install()
    warnings.simplefilter("ignore")
    prob = _make_problem(cls, kw)
    t0 = time.time()
    res = prob.minimize_proximal_gradient(x0, return_all=return_all, **opts)
    out = dict(x=np.asarray(res.x, dtype=np.float64),
               fun=np.asarray(res.fun, dtype=np.float64),
               nit=int(res.nit), success=bool(res.success), time=time.time() - t0)
    if return_all:
        out["allerrs"] = np.asarray(res.allerrs, dtype=np.float64)
        out["allfuns"] = np.asarray(res.allfuns, dtype=np.float64)
    return out


def gen_problem_cases(only, jobs, overwrite):
    specs = problem_specs()
    tasks = []
    for pname, (cls, kw, low, high, n) in specs.items():
        algos = HEAVY.get(pname, DEFAULT_ALGOS)
        algos = list(algos) + [f"fista_ab{i}" for i in SWEEP_PROBLEMS.get(pname, [])]
        rng = np.random.RandomState(sum(map(ord, pname)))
        n_features = _make_problem(cls, kw).n_features
        X0 = rng.uniform(low=low, high=high, size=(n, n_features))
        for algo in algos:
            case = f"{pname}__{algo}"
            if only and only not in case:
                continue
            path = os.path.join(HERE, case + ".npz")
            if os.path.exists(path) and not overwrite:
                continue
            tasks.append((case, path, cls, kw, X0, algo))
    for case, path, cls, kw, X0, algo in tasks:
        opts = dict(ALGOS[algo])
        # benchmarks/benchmark.py:310-311 uses tol_internal=1e-11, max_iter=1e8; keep
        # API defaults otherwise.
        opts.update(tol_internal=1e-11, max_iter=100000000)
        if _make_problem(cls, kw).n_objectives >= 3:
            # trust-constr can take minutes on ONE subproblem with the default
            # max_iter_internal = 100000 (TRIDIA FISTA: 40 outer iterations did not finish in
            # 10 min), so these cases bound both loops through the reference's own options
            opts.update(max_iter=40, max_iter_internal=1000)
        t0 = time.time()
        # first start also records the per-iteration trace
        outs = Parallel(n_jobs=jobs)(
            delayed(_solve_one)(cls, kw, X0[i], opts, i == 0) for i in range(len(X0)))
        save = dict(
            problem=cls, x0=X0,
            x=np.stack([o["x"] for o in outs]),
            fun=np.stack([o["fun"] for o in outs]),
            nit=np.array([o["nit"] for o in outs]),
            success=np.array([o["success"] for o in outs]),
            ref_seconds=np.array([o["time"] for o in outs]),
            allerrs0=outs[0]["allerrs"], allfuns0=outs[0]["allfuns"],
            opt_keys=np.array(sorted(opts)), opt_vals=np.array(
                [repr(opts[k]) for k in sorted(opts)]),
        )
        for k, v in kw.items():
            if k == "bounds":
                save["kw_bounds"] = np.array(v, dtype=np.float64)
            else:
                save["kw_" + k] = np.asarray(v, dtype=np.float64)
        np.savez(path, **save)
        print(f"[golden] {case}: nit={save['nit'].tolist()} "
              f"{time.time() - t0:.1f}s", flush=True)


def _lasso_closures(A, b, l1_ratio, scale):
    """The closures of tests/test_proximal_gradient.py (f = scale*||Ax-b||^2 etc.)."""
    def f(x):
        return np.linalg.norm(A @ x - b) ** 2 * scale

    def g(x):
        return l1_ratio * np.linalg.norm(x, ord=1)

    def jac_f(x):
        return A.T @ (A @ x - b) * (2 * scale)

    def prox_wsum_g(weight, x):
        return np.sign(x) * np.maximum(np.abs(x) - l1_ratio * weight, 0)

    return f, g, jac_f, prox_wsum_g


def gen_lasso_cases(only, overwrite):
    """Single-objective LASSO fixtures (tests/test_proximal_gradient.py:46-111 toy
    problems and its build_dataset() ill-posed regression, plus a fixed-step run in
    the style of examples/cameraman.ipynb: lr=1/L, decay_rate=1)."""
    from zfista import minimize_proximal_gradient

    path = os.path.join(HERE, "lasso_single.npz")
    if only and only not in "lasso_single":
        return
    if os.path.exists(path) and not overwrite:
        return
    warnings.simplefilter("ignore")
    save = {}
    # toy: A = [[-1],[0],[1]], b = [-1,0,1], f = ||Ax-b||^2/6
    A = np.array([[-1.0], [0.0], [1.0]])
    b = np.array([-1.0, 0.0, 1.0])
    x0 = np.array([0.3745401188473625])
    rows = []
    for l1 in [1e-8, 0.1, 0.5, 1.0]:
        for nest in (False, True):
            f, g, jac_f, prox = _lasso_closures(A, b, l1, 1 / 6)
            r = minimize_proximal_gradient(f, g, jac_f, prox, x0, nesterov=nest)
            rows.append([l1, float(nest), float(r.x[0]), float(r.fun), r.nit])
    save["toy_A"], save["toy_b"], save["toy_x0"] = A, b, x0
    save["toy_rows"] = np.array(rows)
    # build_dataset(): 50 x 200, 10 informative
    rs = np.random.RandomState(0)
    w = rs.randn(200)
    w[10:] = 0.0
    X = rs.randn(50, 200)
    y = X @ w
    x0 = np.zeros(200)
    L = 2 * (1 / (2 * 50)) * np.linalg.norm(X, 2) ** 2
    for tag, opts in {
        "bt_ista": dict(nesterov=False),
        "bt_fista": dict(nesterov=True),
        "fixed_fista": dict(nesterov=True, lr=1 / L, decay_rate=1),
        "fixed_fista_ab": dict(nesterov=True, lr=1 / L, decay_rate=1,
                               nesterov_ratio=(0.5, 1 / 16)),
    }.items():
        f, g, jac_f, prox = _lasso_closures(X, y, 0.1, 1 / 100)
        r = minimize_proximal_gradient(f, g, jac_f, prox, x0, return_all=True,
                                       max_iter=20000, **opts)
        save[f"ds_{tag}_x"] = np.asarray(r.x)
        save[f"ds_{tag}_fun"] = np.asarray(r.fun)
        save[f"ds_{tag}_nit"] = np.array(r.nit)
        save[f"ds_{tag}_allerrs"] = np.asarray(r.allerrs)
        save[f"ds_{tag}_allfuns"] = np.asarray(r.allfuns)
        print(f"[golden] lasso ds_{tag}: nit={r.nit}", flush=True)
    save["ds_A"], save["ds_b"], save["ds_x0"], save["ds_L"] = X, y, x0, np.array(L)
    save["ds_l1"], save["ds_scale"] = np.array(0.1), np.array(1 / 100)
    np.savez(path, **save)


def gen_lasso_sweep_cases(only, overwrite):
    """Many runs sharing one A: the dense LASSO of build_dataset() solved by the UNMODIFIED
    reference once per (a, b) pair of the momentum grid (the joblib fan-out of
    examples/PGM_experiment_with_various_a_b.ipynb run()), from per-run starting points, with
    and without the line search.  Parity target of the lockstep device path
    (csrc/zf_lasso_multi.cu)."""
    from zfista import minimize_proximal_gradient

    path = os.path.join(HERE, "lasso_ab_sweep.npz")
    if only and only not in "lasso_ab_sweep":
        return
    if os.path.exists(path) and not overwrite:
        return
    warnings.simplefilter("ignore")
    grid = [(0.0, 0.0), (0.0, 1 / 8), (0.0, 1 / 4), (1 / 6, 1 / 144), (1 / 6, 37 / 288),
            (1 / 6, 1 / 4), (1 / 4, 1 / 64), (1 / 4, 17 / 128), (1 / 4, 1 / 4), (1 / 2, 1 / 16),
            (1 / 2, 5 / 32), (1 / 2, 1 / 4), (3 / 4, 9 / 64), (3 / 4, 25 / 128), (3 / 4, 1 / 4)]
    rs = np.random.RandomState(0)
    w = rs.randn(200)
    w[10:] = 0.0
    X = rs.randn(50, 200)
    y = X @ w
    X0 = np.random.RandomState(1).randn(len(grid), 200) * 0.1
    L = 2 * (1 / (2 * 50)) * np.linalg.norm(X, 2) ** 2
    save = {"A": X, "b": y, "X0": X0, "grid": np.array(grid), "L": np.array(L),
            "l1": np.array(0.1), "scale": np.array(1 / 100)}
    f, g, jac_f, prox = _lasso_closures(X, y, 0.1, 1 / 100)
    for tag, opts in {"bt": dict(), "fixed": dict(lr=1 / L, decay_rate=1)}.items():
        xs, funs, nits = [], [], []
        for k, ab in enumerate(grid):
            r = minimize_proximal_gradient(f, g, jac_f, prox, X0[k], nesterov=True,
                                           nesterov_ratio=ab, max_iter=20000, **opts)
            xs.append(np.asarray(r.x))
            funs.append(float(r.fun))
            nits.append(r.nit)
        save[f"{tag}_x"], save[f"{tag}_fun"], save[f"{tag}_nit"] = (np.array(xs), np.array(funs),
                                                                  np.array(nits))
        print(f"[golden] lasso (a, b) sweep {tag}: nit={nits}", flush=True)
    np.savez(path, **save)


def gen_deblur_cases(only, overwrite):
    """Cameraman-notebook workload at small sizes: the UNMODIFIED reference solver
    (zfista.minimize_proximal_gradient) driven by the notebook's closures as restated in
    oracle/deblur_oracle.py (fixed step lr = 1/L, decay_rate = 1, several (a, b) pairs, and
    one backtracking run)."""
    from zfista import minimize_proximal_gradient

    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import deblur_oracle as do

    path = os.path.join(HERE, "deblur.npz")
    if only and only not in "deblur":
        return
    if os.path.exists(path) and not overwrite:
        return
    warnings.simplefilter("ignore")
    save = {}
    for tag, (h, w, ks, sigma, l1) in {
        "s32": (32, 32, 9, 4.0, 2e-5),
        "s48x64": (48, 64, 5, 1.5, 1e-4),
    }.items():
        kernel = do.gaussian_kernel(ks, sigma)
        kernel = kernel / kernel.sum()
        img, obs, _ = do.synthetic_scene(h, w, seed=h + w, kernel=kernel)
        f, g, jac_f, prox = do.closures(obs, kernel, l1)
        x0 = do.dwt_array(obs)
        L = do.lipschitz(kernel)
        save[f"{tag}_observed"], save[f"{tag}_kernel"] = obs, kernel
        save[f"{tag}_l1"], save[f"{tag}_L"], save[f"{tag}_x0"] = np.array(l1), np.array(L), x0
        pairs = [AB_GRID[i] for i in (0, 2, 4, 8, 14)]
        save[f"{tag}_pairs"] = np.array(pairs)
        for i, ab in enumerate(pairs):
            r = minimize_proximal_gradient(f, g, jac_f, prox, x0, lr=1 / L, decay_rate=1,
                                           nesterov=True, nesterov_ratio=ab, return_all=True,
                                           max_iter=400, tol=1e-5)
            save[f"{tag}_fixed{i}_x"] = np.asarray(r.x)
            save[f"{tag}_fixed{i}_fun"] = np.asarray(r.fun).reshape(-1)
            save[f"{tag}_fixed{i}_nit"] = np.array(r.nit)
            save[f"{tag}_fixed{i}_success"] = np.array(bool(r.success))
            save[f"{tag}_fixed{i}_allerrs"] = np.asarray(r.allerrs)
            save[f"{tag}_fixed{i}_allfuns"] = np.asarray(r.allfuns).reshape(-1)
            print(f"[golden] deblur {tag} fixed ab={ab}: nit={r.nit} ok={r.success}", flush=True)
        for name, opts in {"bt_fista": dict(nesterov=True), "bt_ista": dict(nesterov=False)}.items():
            r = minimize_proximal_gradient(f, g, jac_f, prox, x0, return_all=True, max_iter=150,
                                           **opts)
            save[f"{tag}_{name}_x"] = np.asarray(r.x)
            save[f"{tag}_{name}_fun"] = np.asarray(r.fun).reshape(-1)
            save[f"{tag}_{name}_nit"] = np.array(r.nit)
            save[f"{tag}_{name}_allerrs"] = np.asarray(r.allerrs)
            save[f"{tag}_{name}_allfuns"] = np.asarray(r.allfuns).reshape(-1)
            print(f"[golden] deblur {tag} {name}: nit={r.nit} ok={r.success}", flush=True)
        # closure values at random points (pins W, W^T, R as used by the solver)
        rng = np.random.RandomState(3)
        X = rng.standard_normal((3, h * w)) * 0.1
        save[f"{tag}_evalX"] = X
        save[f"{tag}_evalf"] = np.array([f(x)[0] for x in X])
        save[f"{tag}_evaljac"] = np.array([jac_f(x)[0] for x in X])
    np.savez(path, **save)


def gen_subproblem_cases(only, overwrite):
    """Direct fixtures of zfista.proximal_gradient._solve_subproblem (35-209):
    (yk, xk_old, lr) -> (weight, x, fun) for m = 2 (bounded Brent) and m >= 3
    (trust-constr)."""
    from zfista.proximal_gradient import _solve_subproblem

    path = os.path.join(HERE, "subproblem.npz")
    if only and only not in "subproblem":
        return
    if os.path.exists(path) and not overwrite:
        return
    warnings.simplefilter("ignore")
    specs = problem_specs()
    save = {}
    names = ["JOS1_n5", "JOS1_n5_l1", "JOS1_n50_l1", "SD", "ZDT1_n50", "TOI4_l1",
             "TRIDIA", "TRIDIA_l1", "LFR1_n30", "FDS_n10", "FDS_n10_l1", "FDS_n10_box",
             "FDS_n100_l1"]
    for pname in names:
        cls, kw, low, high, _ = specs[pname]
        prob = _make_problem(cls, kw)
        rng = np.random.RandomState(7 + sum(map(ord, pname)))
        n, m = prob.n_features, prob.n_objectives
        K = 12
        Y = rng.uniform(low, high, size=(K, n))
        XO = Y + 0.05 * rng.standard_normal((K, n)) * (np.arange(K)[:, None] % 3 > 0)
        if prob.bounds is not None:
            XO = np.clip(XO, prob.bounds[0], prob.bounds[1])
            Y = np.clip(Y, prob.bounds[0], prob.bounds[1])
        LR = np.array([1.0, 0.5, 0.25, 0.125, 1 / 64, 1 / 1024] * 2)[:K]
        W, X, FUN = np.zeros((K, m)), np.zeros((K, n)), np.zeros(K)
        for k in range(K):
            w0 = np.ones(m) / m
            r = _solve_subproblem(prob.f, prob.g, prob.jac_f, prob.prox_wsum_g,
                                  LR[k], XO[k], Y[k], w0, tol=1e-11, max_iter=100000,
                                  deprecated=(k % 4 == 3))
            W[k], X[k], FUN[k] = r.weight, np.asarray(r.x), r.fun
        save[pname + "__y"], save[pname + "__xold"], save[pname + "__lr"] = Y, XO, LR
        save[pname + "__w"], save[pname + "__x"], save[pname + "__fun"] = W, X, FUN
        print(f"[golden] subproblem {pname}", flush=True)
    np.savez(path, **save)


def gen_problem_eval_cases(only, overwrite):
    """f / jac_f / g / prox_wsum_g values of every class in zfista/problems.py at
    random points (pins the device functors)."""
    path = os.path.join(HERE, "problem_eval.npz")
    if only and only not in "problem_eval":
        return
    if os.path.exists(path) and not overwrite:
        return
    specs = problem_specs()
    save = {}
    for pname, (cls, kw, low, high, _) in specs.items():
        prob = _make_problem(cls, kw)
        rng = np.random.RandomState(11 + sum(map(ord, pname)))
        n, m = prob.n_features, prob.n_objectives
        K = 6
        X = rng.uniform(low, high, size=(K, n))
        Wt = rng.uniform(0, 1, size=(K, m))
        F = np.stack([prob.f(x) for x in X])
        G = np.stack([prob.g(x) for x in X])
        J = np.stack([prob.jac_f(x) for x in X])
        P = np.stack([np.asarray(prob.prox_wsum_g(w, x)) for w, x in zip(Wt, X)])
        save[pname + "__X"], save[pname + "__W"] = X, Wt
        save[pname + "__f"], save[pname + "__g"] = F, G
        save[pname + "__jac"], save[pname + "__prox"] = J, P
    np.savez(path, **save)
    print("[golden] problem_eval", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--jobs", type=int, default=8)
    ap.add_argument("--overwrite", action="store_true")
    a = ap.parse_args()
    gen_problem_eval_cases(a.only, a.overwrite)
    gen_lasso_cases(a.only, a.overwrite)
    gen_lasso_sweep_cases(a.only, a.overwrite)
    gen_deblur_cases(a.only, a.overwrite)
    gen_subproblem_cases(a.only, a.overwrite)
    gen_problem_cases(a.only, a.jobs, a.overwrite)
