"""Shared helpers for the tests: load golden fixtures and build, from the same
constructor arguments, (a) the oracle's ProblemSpec and (b) the zfista_b200 Problem."""
from __future__ import annotations

import ast
import glob
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

AB_GRID = [
    (0.0, 0.0), (0.0, 1 / 8), (0.0, 1 / 4),
    (1 / 6, 1 / 144), (1 / 6, 37 / 288), (1 / 6, 1 / 4),
    (1 / 4, 1 / 64), (1 / 4, 17 / 128), (1 / 4, 1 / 4),
    (1 / 2, 1 / 16), (1 / 2, 5 / 32), (1 / 2, 1 / 4),
    (3 / 4, 9 / 64), (3 / 4, 25 / 128), (3 / 4, 1 / 4),
]


def golden_cases():
    """Names of the per-problem full-solve fixtures present (<Problem>__<algo>)."""
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*__*.npz")))


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def case_kwargs(d):
    """Constructor kwargs stored in a fixture (kw_* arrays)."""
    kw = {}
    for k in d.files:
        if not k.startswith("kw_"):
            continue
        v = d[k]
        key = k[3:]
        if key == "bounds":
            kw["bounds"] = (float(v[0]), float(v[1]))
        elif key in ("n_features", "n_objectives"):
            kw[key] = int(v)
        else:
            kw[key] = np.array(v, dtype=np.float64)
    return kw


def case_options(d):
    """Solver options a fixture was generated with."""
    return {str(k): ast.literal_eval(str(v)) for k, v in zip(d["opt_keys"], d["opt_vals"])}


def oracle_spec(cls, kw):
    from oracle import zfista_oracle as zo

    kw = dict(kw)
    bounds = kw.pop("bounds", None)
    return zo.make_spec(cls, bounds=bounds, **kw)


def device_problem(cls, kw):
    import zfista_b200.problems as zp

    return getattr(zp, cls)(**kw)


# name -> (class, kwargs) for the subproblem / problem_eval fixtures (same table as
# tests/golden/make_golden.py:problem_specs, restated so tests do not import the generator)
def fixture_problems():
    def l1(n, m):
        return dict(l1_ratios=(np.arange(m) + 1) / n, l1_shifts=np.arange(m))

    return {
        "JOS1_n5": ("JOS1", dict(n_features=5)),
        "JOS1_n50": ("JOS1", dict(n_features=50)),
        "JOS1_n200": ("JOS1", dict(n_features=200)),
        "JOS1_n5_l1": ("JOS1", dict(n_features=5, **l1(5, 2))),
        "JOS1_n50_l1": ("JOS1", dict(n_features=50, **l1(50, 2))),
        "JOS1_n50_l1nb": ("JOS1", dict(n_features=50, l1_ratios=(1 / 50, 1 / 100),
                                       l1_shifts=(0, 1))),
        "SD": ("SD", dict()),
        "ZDT1_n50": ("ZDT1", dict(n_features=50)),
        "TOI4": ("TOI4", dict()),
        "TOI4_l1": ("TOI4", l1(4, 2)),
        "TRIDIA": ("TRIDIA", dict()),
        "TRIDIA_l1": ("TRIDIA", l1(3, 3)),
        "LFR1_n30": ("LinearFunctionRank1", dict(n_features=30)),
        "FDS_n5": ("FDS", dict(n_features=5)),
        "FDS_n10": ("FDS", dict(n_features=10)),
        "FDS_n10_l1": ("FDS", dict(n_features=10, **l1(10, 3))),
        "FDS_n10_box": ("FDS", dict(n_features=10, bounds=(0, np.inf))),
        "FDS_n100_l1": ("FDS", dict(n_features=100, **l1(100, 3))),
    }


# ---------------------------------------------------------------------------------------
# Rounding sensitivity of the REFERENCE itself.
#
# For two objectives the reference finds the dual weight with scipy's bounded Brent, whose
# resolution is sqrt(eps)*|w| ~ 1.5e-8 and whose path depends on comparisons of function
# values that differ in the last bit.  On problems with L1 terms (kinks) or flat directions
# (TOI4) the outer FISTA loop amplifies that: perturbing the reference's own dual function
# by ONE ulp changes its iteration count by +-10 % and its final x by up to ~1e-3 (see
# tests/test_oracle_golden.py::test_reference_is_rounding_sensitive_on_l1_cases).  The
# reference's numbers on such cases are therefore one draw from an envelope (they also
# depend on the BLAS build numpy links to); a different summation order -- any GPU -- is
# another draw.  `oracle_noise_envelope` measures that envelope with the oracle so that a
# test can require "the device is as close to the reference as the reference is to itself".
# ---------------------------------------------------------------------------------------
def noisy_brent_subproblem(noise, seed):
    """oracle subproblem (m = 2, restated Brent) whose dual value carries `noise` relative
    rounding noise; noise = 0 reproduces the reference bit for bit."""
    from oracle import dual_model as dm
    from oracle import zfista_oracle as zo

    rng = np.random.RandomState(seed)

    def sub(spec, lr, x_prev, y, w_init, tol=1e-12, max_iter=1000, deprecated=False):
        fy = zo.f(spec, y)
        F_prev = zo.f(spec, x_prev) + zo.g(spec, x_prev)
        Jy = zo.jac_f(spec, y)

        def neg_dual(ws):
            w = np.array([ws, 1 - ws])
            wj = w @ Jy
            v = y - lr * wj
            p = zo.prox_wsum_g(spec, lr * w, v)
            val = (-np.inner(w, zo.g(spec, p)) - np.linalg.norm(p - v) ** 2 / 2 / lr
                   + lr / 2 * np.linalg.norm(wj) ** 2)
            if not deprecated:
                val += np.inner(w, F_prev - fy)
            return val * (1 + noise * rng.uniform(-1, 1))

        xf, fx, nfev = dm.fmin_bounded(neg_dual, 0.0, 1.0, xatol=tol, maxfun=max_iter)
        weight = np.array([xf, 1 - xf])
        x = zo.prox_wsum_g(spec, lr * weight, y - lr * weight @ Jy)
        return x, -fx, weight, nfev

    return sub


def oracle_noise_envelope(spec, X0, x_ref, fun_ref, nit_ref, opts, n_starts=4, seeds=(0, 1, 2),
                          noise=1.1e-16):
    """max deviation of the (1-ulp perturbed) oracle from the reference's stored result over
    the first `n_starts` starts: dict(dx, dF (relative), dnit)."""
    import warnings

    from oracle import zfista_oracle as zo

    dx = dF = 0.0
    dnit = 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(min(n_starts, len(X0))):
            for s in seeds:
                r = zo.minimize_proximal_gradient(
                    spec, X0[i], subproblem=noisy_brent_subproblem(noise, s), **opts)
                dx = max(dx, float(np.max(np.abs(r["x"] - x_ref[i]))))
                dF = max(dF, float(np.max(np.abs(r["fun"] - fun_ref[i])
                                          / np.maximum(1.0, np.abs(fun_ref[i])))))
                dnit = max(dnit, abs(int(r["nit"]) - int(nit_ref[i])))
    return dict(dx=dx, dF=dF, dnit=dnit)


# ---------------------------------------------------------------------------------------
# m >= 3: converged fixtures of the unmodified reference (tests/golden/make_golden_converged.py)
# and the CPU statement of the device algorithm (oracle DeviceModel)
# ---------------------------------------------------------------------------------------
def converged_cases():
    return sorted(os.path.basename(p)[:-4]
                  for p in glob.glob(os.path.join(GOLDEN, "converged", "*__*.npz")))


def load_converged(name):
    return np.load(os.path.join(GOLDEN, "converged", name + ".npz"), allow_pickle=False)


def converged_trace(d, i):
    """(allerrs, allfuns) of start i of a converged fixture (ragged traces, stored flat)."""
    o = d["trace_offsets"]
    return d["allerrs"][o[i]:o[i + 1]], d["allfuns"][o[i] + i:o[i + 1] + i + 1]


def device_model_solve(spec, x0, opts, newton_for_two=False):
    import warnings

    from oracle import zfista_oracle as zo

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return zo.minimize_proximal_gradient(
            spec, x0, subproblem=zo.DeviceModel(newton_for_two=newton_for_two), **opts)


def device_model_envelope(spec, x0, opts, ref=None, seeds=(0, 1, 2, 3)):
    """How far the CPU statement of the device algorithm moves when x0 is perturbed by ONE ulp
    (a different summation order -- numpy pairwise on the CPU, lane-strided + butterfly on the
    GPU -- is a perturbation of at least that size at every operation).  Returns the unperturbed
    result and dict(dnit, dx, dF) = the largest deviation over the seeds."""
    r = device_model_solve(spec, x0, opts) if ref is None else ref
    dnit, dx, dF = 0, 0.0, 0.0
    for s in seeds:
        sg = np.sign(np.random.RandomState(s).uniform(-1, 1, x0.shape[0]))
        rp = device_model_solve(spec, x0 * (1 + 1.1e-16 * sg), opts)
        dnit = max(dnit, abs(int(rp["nit"]) - int(r["nit"])))
        dx = max(dx, float(np.max(np.abs(rp["x"] - r["x"]))))
        dF = max(dF, float(np.max(np.abs(rp["fun"] - r["fun"]) / np.maximum(1.0, np.abs(r["fun"])))))
    return r, dict(dnit=dnit, dx=dx, dF=dF)
