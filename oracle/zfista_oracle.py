"""CPU oracle for the zfista proximal-gradient hot path.  TEST INFRASTRUCTURE ONLY.

This module is a numpy restatement of the reference algorithm
(/root/reference/zfista/proximal_gradient.py and zfista/problems.py) used as the
checker for the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product (``zfista_b200``) never does.

PARITY PINNED: every function here is checked against fixtures produced by running
the unmodified reference in the build container (tests/golden/make_golden.py, via
tests/golden/refshim.py) -- see tests/test_oracle_golden.py.

Third-party pieces and how they are handled
-------------------------------------------
* scipy (present in this image, 1.18.1): the reference's inner solvers are
  ``scipy.optimize.minimize_scalar(bounds=(0, 1))`` for two objectives
  (proximal_gradient.py:184-188) and ``scipy.optimize.minimize(method=
  "trust-constr", hess=BFGS())`` for three or more (proximal_gradient.py:193-202).
  The oracle calls the same scipy entry points with the same options, because
  the result of the reference IS whatever those routines return.
* jaxopt (absent; pyproject.toml dependency, unpinned, public 0.8.x): only
  ``prox_lasso`` and ``projection_box`` are used (problems.py:8-9, 128-137); their
  published definitions, sign(x)*max(|x|-t,0) and clip(x, lo, hi), are restated in
  :func:`soft_threshold` and :func:`prox_wsum_g`.

Layout differs from the reference on purpose: problems are plain ``ProblemSpec``
records dispatched by name (the same descriptor the C-ABI takes), and the solver
is one function over that record.
"""
from __future__ import annotations

import time
import warnings
from dataclasses import dataclass, field

import numpy as np
from scipy.optimize import BFGS, Bounds, LinearConstraint, minimize, minimize_scalar

SQRT2 = np.sqrt(2)

# problem ids shared with include/zfista_b200.h
PROBLEM_IDS = {
    "JOS1": 0, "SD": 1, "FDS": 2, "ZDT1": 3, "TOI4": 4, "TRIDIA": 5,
    "LinearFunctionRank1": 6, "LeastSquaresL1": 7,
}
FIXED_SHAPES = {"SD": (4, 2), "TOI4": (4, 2), "TRIDIA": (3, 3)}
DEFAULT_N = {"JOS1": 5, "FDS": 10, "ZDT1": 30, "LinearFunctionRank1": 10}
N_OBJECTIVES = {"JOS1": 2, "SD": 2, "FDS": 3, "ZDT1": 2, "TOI4": 2, "TRIDIA": 3}


@dataclass
class ProblemSpec:
    """Descriptor of one built-in problem class (zfista/problems.py:25-150)."""

    kind: str
    n_features: int
    n_objectives: int
    l1_ratios: np.ndarray | None = None
    l1_shifts: np.ndarray | None = None
    lower: object = None          # scalar, array (n,) or None
    upper: object = None
    # LeastSquaresL1 only: f(x) = scale * ||A x - b||^2, g(x) = l1 * ||x||_1
    A: np.ndarray | None = None
    b: np.ndarray | None = None
    scale: float = 1.0
    l1: float = 0.0
    extra: dict = field(default_factory=dict)

    @property
    def has_bounds(self):
        return self.lower is not None


def make_spec(kind, n_features=None, n_objectives=None, l1_ratios=None, l1_shifts=None,
              bounds=None):
    """Build a spec with the reference constructors' defaults
    (problems.py:177-191, 238-245, 294-310, 365-366, 415-428, 479-492, 540-556)."""
    if kind in FIXED_SHAPES:
        n, m = FIXED_SHAPES[kind]
    else:
        n = DEFAULT_N[kind] if n_features is None else int(n_features)
        if kind == "LinearFunctionRank1":
            m = 4 if n_objectives is None else int(n_objectives)
        else:
            m = N_OBJECTIVES[kind]
    if kind in ("SD", "ZDT1"):
        bounds = (1e-6, np.inf)          # problems.py:244, 366
        l1_ratios = l1_shifts = None
    spec = ProblemSpec(kind, n, m)
    if l1_ratios is not None:
        spec.l1_ratios = np.array(l1_ratios, dtype=np.float64)
    spec.l1_shifts = (np.zeros(m) if l1_shifts is None
                      else np.array(l1_shifts, dtype=np.float64))
    if bounds is not None:
        spec.lower, spec.upper = bounds
    return spec


def make_least_squares_l1(A, b, l1, scale=1.0):
    """Single-objective  scale*||Ax-b||^2 + l1*||x||_1  (the closures of
    tests/test_proximal_gradient.py:49-63 with scale=1/6, and of
    examples/cameraman.ipynb with scale=1)."""
    A = np.asarray(A, dtype=np.float64)
    spec = ProblemSpec("LeastSquaresL1", A.shape[1], 1, A=A,
                       b=np.asarray(b, dtype=np.float64), scale=float(scale), l1=float(l1))
    return spec


# --------------------------------------------------------------------------
# f, jac_f  (problems.py:193-205, 247-264, 312-328, 368-386, 430-448, 494-514, 558-578)
# --------------------------------------------------------------------------
def f(spec: ProblemSpec, x):
    k, n = spec.kind, spec.n_features
    if k == "JOS1":
        return np.array([np.linalg.norm(x) ** 2 / n, np.linalg.norm(x - 2) ** 2 / n])
    if k == "SD":
        return np.array([
            2 * x[0] + SQRT2 * x[1] + SQRT2 * x[2] + x[3],
            2 / x[0] + 2 * SQRT2 / x[1] + 2 * SQRT2 / x[2] + 2 / x[3],
        ])
    if k == "FDS":
        idx = np.arange(n) + 1
        conv = idx * idx[::-1]
        return np.array([
            np.inner(idx, (x - idx) ** 4) / n ** 2,
            np.exp(x.sum() / n) + np.linalg.norm(x) ** 2,
            np.inner(conv, np.exp(-x)) / (n * (n + 1)),
        ])
    if k == "ZDT1":
        h = 1 + 9 / (n - 1) * np.sum(x[1:])
        return np.array([x[0], h * (1 - np.sqrt(x[0] / h))])
    if k == "TOI4":
        return np.array([
            x[0] ** 2 + x[1] ** 2 + 1,
            0.5 * ((x[0] - x[1]) ** 2 + (x[2] - x[3]) ** 2) + 1,
        ])
    if k == "TRIDIA":
        return np.array([
            (2 * x[0] - 1) ** 2,
            2 * (2 * x[0] - x[1]) ** 2,
            3 * (2 * x[1] - x[2]) ** 2,
        ])
    if k == "LinearFunctionRank1":
        io = np.arange(1, spec.n_objectives + 1)
        jf = np.arange(1, n + 1)
        return (io * np.inner(jf, x) - 1) ** 2
    if k == "LeastSquaresL1":
        return np.linalg.norm(spec.A @ x - spec.b) ** 2 * spec.scale
    if k == "Closures":        # caller-supplied numpy closures (oracle/deblur_oracle.py)
        return spec.extra["f"](x)
    raise ValueError(k)


def jac_f(spec: ProblemSpec, x):
    k, n = spec.kind, spec.n_features
    if k == "JOS1":
        return np.vstack((2 * x / n, 2 * (x - 2) / n))
    if k == "SD":
        return np.vstack((
            np.array([2, SQRT2, SQRT2, 1]),
            np.array([-2 / x[0] ** 2, -2 * SQRT2 / x[1] ** 2,
                      -2 * SQRT2 / x[2] ** 2, -2 / x[3] ** 2]),
        ))
    if k == "FDS":
        idx = np.arange(n) + 1
        conv = idx * idx[::-1]
        return np.vstack((
            4 / n ** 2 * idx * (x - idx) ** 3,
            np.exp(x.sum() / n) / n + 2 * x,
            -conv * np.exp(-x) / (n * (n + 1)),
        ))
    if k == "ZDT1":
        h = 1 + 9 / (n - 1) * np.sum(x[1:])
        j1 = np.zeros(n)
        j1[0] = 1
        j2 = np.full(n, 9 * (2 - np.sqrt(x[0] / h)) / 2 / (n - 1))
        j2[0] = -np.sqrt(h / x[0]) / 2
        return np.vstack((j1, j2))
    if k == "TOI4":
        j1 = np.zeros(n)
        j1[0], j1[1] = 2 * x[0], 2 * x[1]
        j2 = np.zeros(n)
        j2[0] = x[0] - x[1]
        j2[1] = -j2[0]
        j2[2] = x[2] - x[3]
        j2[3] = -j2[2]
        return np.vstack((j1, j2))
    if k == "TRIDIA":
        return np.array([
            [8 * x[0] - 4, 0, 0],
            [16 * x[0] - 8 * x[1], 4 * x[1] - 8 * x[0], 0],
            [0, 24 * x[1] - 12 * x[2], 6 * x[2] - 12 * x[1]],
        ])
    if k == "LinearFunctionRank1":
        io = np.arange(1, spec.n_objectives + 1)
        jf = np.arange(1, n + 1)
        return 2 * io[:, None] * jf * (io[:, None] * np.inner(jf, x) - 1)
    if k == "LeastSquaresL1":
        return spec.A.T @ (spec.A @ x - spec.b) * (2 * spec.scale)
    if k == "Closures":
        return spec.extra["jac_f"](x)
    raise ValueError(k)


# --------------------------------------------------------------------------
# g, prox_wsum_g  (problems.py:101-138)
# --------------------------------------------------------------------------
def soft_threshold(x, t):
    """jaxopt.prox.prox_lasso(x, t): sign(x) * max(|x| - t, 0)."""
    return np.sign(x) * np.maximum(np.abs(x) - t, 0.0)


def g(spec: ProblemSpec, x):
    if spec.kind == "LeastSquaresL1":
        return spec.l1 * np.linalg.norm(x, ord=1)
    if spec.kind == "Closures":
        return spec.extra["g"](x)
    m = spec.n_objectives
    if spec.has_bounds:
        if (x < spec.lower).any() or (x > spec.upper).any():
            return np.full(m, np.inf)
    if spec.l1_ratios is not None:
        return spec.l1_ratios * np.linalg.norm(x - spec.l1_shifts.reshape(-1, 1), ord=1,
                                               axis=1)
    return np.zeros(m)


def prox_wsum_g(spec: ProblemSpec, weight, x):
    if spec.kind == "LeastSquaresL1":
        return soft_threshold(x, spec.l1 * weight)
    if spec.kind == "Closures":
        return spec.extra["prox_wsum_g"](weight, x)
    if spec.l1_ratios is not None:
        coef = weight * spec.l1_ratios
        s = spec.l1_shifts
        # problems.py:128-135 (note the first stage adds and subtracts shift 0)
        x = soft_threshold(x + np.sum(coef[1:]) - s[0] + s[0], coef[0])
        for i in range(1, spec.n_objectives):
            x = soft_threshold(x - coef[i] - s[i], coef[i]) + s[i]
    if spec.has_bounds:
        x = np.clip(x, spec.lower, spec.upper)
    return x


# --------------------------------------------------------------------------
# the solver  (proximal_gradient.py:35-209, 212-308, 311-555)
# --------------------------------------------------------------------------
class BacktrackingFailed(RuntimeError):
    pass


def solve_subproblem(spec, lr, x_prev, y, w_init, tol=1e-12, max_iter=1000,
                     deprecated=False):
    """One proximal subproblem.  Returns (x, fun, weight, inner_iterations)."""
    fy = f(spec, y)
    F_prev = f(spec, x_prev) + g(spec, x_prev)
    Jy = jac_f(spec, y)
    m = spec.n_objectives
    if m == 1:
        x = prox_wsum_g(spec, lr, y - lr * Jy.flatten())
        fun = float(Jy @ (x - y) + g(spec, x) + np.linalg.norm(x - y) ** 2 / 2 / lr)
        if not deprecated:
            fun += fy - F_prev
        return x, fun, None, 1

    def neg_dual(w):
        wj = w @ Jy
        v = y - lr * wj
        p = prox_wsum_g(spec, lr * w, v)
        gp = g(spec, p)
        val = (-np.inner(w, gp) - np.linalg.norm(p - v) ** 2 / 2 / lr
               + lr / 2 * np.linalg.norm(wj) ** 2)
        grad = -gp - Jy @ (p - y)
        if not deprecated:
            val += np.inner(w, F_prev - fy)
            grad += F_prev - fy
        return val, grad

    if m == 2:
        res = minimize_scalar(lambda w: neg_dual(np.array([w, 1 - w]))[0], bounds=(0, 1),
                              options={"maxiter": max_iter, "xatol": tol})
        weight = np.array([res.x, 1 - res.x])
    else:
        res = minimize(fun=neg_dual, x0=w_init, method="trust-constr", jac=True,
                       hess=BFGS(), bounds=Bounds(lb=0, ub=np.inf),
                       constraints=LinearConstraint(np.ones(m), lb=1, ub=1),
                       options={"gtol": tol, "xtol": tol, "barrier_tol": tol,
                                "maxiter": max_iter})
        weight = res.x
    if not res.success:
        warnings.warn(str(res.message), stacklevel=2)
    x = prox_wsum_g(spec, lr * weight, y - lr * weight @ Jy)
    return x, -res.fun, weight, res.nit


def minimize_proximal_gradient(spec, x0, lr=1, tol=1e-5, tol_internal=1e-12,
                               max_iter=1000000, max_iter_internal=100000,
                               max_backtrack_iter=100, warm_start=False, decay_rate=0.5,
                               nesterov=False, nesterov_ratio=(0, 0.25), return_all=False,
                               deprecated=False, subproblem=solve_subproblem):
    """Oracle for zfista.minimize_proximal_gradient on a ProblemSpec.  Returns a
    dict with the OptimizeResult fields of the reference (x, fun, nit, success,
    status, allvecs/allfuns/allerrs) plus ``lr`` (final step) and ``nfev``."""
    t_start = time.time()
    x0 = np.asarray(x0, dtype=np.float64)
    m = spec.n_objectives
    x_prev = x = y = x0
    w = np.ones(m) / m if m > 1 else None
    t_prev = 1
    F0 = f(spec, x0) + g(spec, x0)
    allvecs, allfuns, allerrs = [x0], [F0], []
    out = dict(success=False, status=0, message="Maximum number of iterations reached")
    nit = 0
    for nit in range(1, max_iter + 1):
        # ---- backtracking line search (proximal_gradient.py:279-308) ----
        F_prev = f(spec, x_prev) + g(spec, x_prev)
        try:
            for _ in range(max_backtrack_iter):
                x, sub_fun, weight, _ = subproblem(
                    spec, lr, x_prev, y, w, tol=tol_internal, max_iter=max_iter_internal,
                    deprecated=deprecated)
                F_x = f(spec, x) + g(spec, x)
                if w is not None and warm_start:
                    w = weight
                if decay_rate == 1:
                    break
                if deprecated:
                    if np.all(f(spec, x) - f(spec, y) <= sub_fun + tol_internal):
                        break
                elif np.all(F_x - F_prev <= sub_fun + tol_internal):
                    break
                lr *= decay_rate
            else:
                raise BacktrackingFailed("Backtracking failed to find a suitable stepsize.")
        except Exception as e:  # proximal_gradient.py:493-509
            return dict(success=False, status=-1, message=f"Error: {e}", x=x_prev,
                        fun=f(spec, x_prev) + g(spec, x_prev), nit=nit - 1, lr=lr,
                        time=time.time() - t_start,
                        allvecs=allvecs if return_all else None,
                        allfuns=allfuns if return_all else None,
                        allerrs=allerrs if return_all else None)
        err = max(abs(x - y))
        if return_all:
            allvecs.append(x)
            allfuns.append(f(spec, x) + g(spec, x))
            allerrs.append(err)
        if err < tol:
            out.update(success=True, status=1, message="Optimization terminated successfully")
            break
        if nesterov:
            a, b = nesterov_ratio
            t_new = np.sqrt(t_prev ** 2 - a * t_prev + b) + 0.5
            y = x + (t_prev - 1) / t_new * (x - x_prev)
            t_prev = t_new
        else:
            y = x
        x_prev = x
    out.update(x=x, fun=f(spec, x) + g(spec, x), nit=nit, lr=lr,
               time=time.time() - t_start,
               allvecs=allvecs if return_all else None,
               allfuns=allfuns if return_all else None,
               allerrs=allerrs if return_all else None)
    return out


# --------------------------------------------------------------------------
# Subproblem with the DEVICE's inner solvers (oracle/dual_model.py) instead of scipy.
# Pass as ``subproblem=`` to minimize_proximal_gradient to get the CPU statement of
# exactly what the CUDA kernel computes: bounded Brent for m = 2 (identical to the
# scipy route up to rounding) and the exact simplex Newton for m >= 3, where scipy's
# trust-constr only reaches the optimum to ~1e-4..1e-7 ("Inner solver").
# --------------------------------------------------------------------------
def solve_subproblem_device_model(spec, lr, x_prev, y, w_init, tol=1e-12, max_iter=1000,
                                  deprecated=False, newton_for_two=False):
    from . import dual_model as dm

    m = spec.n_objectives
    if m == 1:
        return solve_subproblem(spec, lr, x_prev, y, w_init, tol, max_iter, deprecated)
    fy = f(spec, y)
    F_prev = f(spec, x_prev) + g(spec, x_prev)
    Jy = jac_f(spec, y)
    if m == 2 and not newton_for_two:
        def neg_dual(ws):
            w = np.array([ws, 1 - ws])
            wj = w @ Jy
            v = y - lr * wj
            p = prox_wsum_g(spec, lr * w, v)
            val = (-np.inner(w, g(spec, p)) - np.linalg.norm(p - v) ** 2 / 2 / lr
                   + lr / 2 * np.linalg.norm(wj) ** 2)
            if not deprecated:
                val += np.inner(w, F_prev - fy)
            return val

        xf, fx, nfev = dm.fmin_bounded(neg_dual, 0.0, 1.0, xatol=tol, maxfun=max_iter)
        weight = np.array([xf, 1 - xf])
        x = prox_wsum_g(spec, lr * weight, y - lr * weight @ Jy)
        return x, -fx, weight, nfev
    c = np.zeros(m) if deprecated else fy - F_prev
    if spec.kind == "LeastSquaresL1":
        lam, sh = np.full(m, spec.l1), np.zeros(m)
    else:
        lam, sh = spec.l1_ratios, spec.l1_shifts
    weight, D, p, evals = dm.simplex_newton(y, Jy, lr, c, lam, sh, spec.lower, spec.upper,
                                            lambda q: g(spec, q), w0=w_init)
    return p, D, weight, evals


class DeviceModel:
    """``subproblem=DeviceModel()`` -- the CPU statement of what ONE start of
    batched_fista_kernel (zf_batched.cu) does around its inner solver: like
    :func:`solve_subproblem_device_model`, plus the one piece of state the kernel carries from
    subproblem to subproblem: the simplex Newton solver always continues from the previous
    subproblem's weights (``wwarm`` in zf_batched.cu; the reference's ``warm_start`` option only
    matters for the Brent route, which has no initial guess).  One instance per solve.
    ``newton_for_two`` = the kernel's ``dual_solver="newton"`` for two objectives."""

    def __init__(self, newton_for_two=False):
        self.newton_for_two = newton_for_two
        self.w = None
        self.dual_evals = 0

    def __call__(self, spec, lr, x_prev, y, w_init, tol=1e-12, max_iter=1000, deprecated=False):
        m = spec.n_objectives
        newton = m >= 3 or (m == 2 and self.newton_for_two)
        w0 = self.w if (newton and self.w is not None) else w_init
        x, fun, weight, evals = solve_subproblem_device_model(
            spec, lr, x_prev, y, w0, tol, max_iter, deprecated, self.newton_for_two)
        if newton:
            self.w = weight
        self.dual_evals += evals
        return x, fun, weight, evals
