"""``minimize_proximal_gradient`` with the signature and OptimizeResult fields of
/root/reference/zfista/proximal_gradient.py:311-555, executed on the GPU.

The reference takes arbitrary Python callables; the device path takes the *device
functors* of :mod:`zfista_b200.problems`.  Pass the bound methods of one Problem
instance (``p.f, p.g, p.jac_f, p.prox_wsum_g``) exactly as the reference's
``Problem.minimize_proximal_gradient`` does (problems.py:140-150).  Any other callable
raises ``TypeError``: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import time
from dataclasses import dataclass
from warnings import warn

import numpy as np
from scipy.optimize import OptimizeResult

from . import _lib
from .problems import Problem, _as_f64, _ptr

# proximal_gradient.py:19-22
TERMINATION_MESSAGES = {
    0: "The maximum number of iterations is exceeded.",
    1: "Termination condition is satisfied.",
}

DUAL_SOLVERS = {"reference": 0, "newton": 1}
_DEFAULT_TRACE = 4096


def _resolve_problem(f, g, jac_f, prox_wsum_g) -> Problem:
    if isinstance(f, Problem) and g is None and jac_f is None and prox_wsum_g is None:
        return f
    owners = []
    for name, fn in (("f", f), ("g", g), ("jac_f", jac_f), ("prox_wsum_g", prox_wsum_g)):
        owner = getattr(fn, "__self__", None)
        if not isinstance(owner, Problem) or getattr(fn, "__name__", None) != name:
            raise TypeError(
                f"{name} must be the bound method `{name}` of a zfista_b200.problems.Problem "
                "(a device functor). Arbitrary Python callables cannot run on the GPU and "
                "zfista_b200 has no CPU fallback; express the objective as a Problem "
                "(e.g. LeastSquaresL1) instead.")
        owners.append(owner)
    if any(o is not owners[0] for o in owners):
        raise TypeError("f, g, jac_f and prox_wsum_g must belong to the same Problem instance")
    return owners[0]


def _make_options(lr, tol, tol_internal, max_iter, max_iter_internal, max_backtrack_iter,
                  warm_start, decay_rate, nesterov, nesterov_ratio, deprecated, dual_solver,
                  trace_capacity) -> _lib.ZfOptions:
    if dual_solver not in DUAL_SOLVERS:
        raise ValueError(f"dual_solver must be one of {sorted(DUAL_SOLVERS)}")
    o = _lib.default_options()
    o.lr = float(lr)
    o.tol = float(tol)
    o.tol_internal = float(tol_internal)
    o.max_iter = int(max_iter)
    o.max_iter_internal = int(min(max_iter_internal, 2**31 - 1))
    o.max_backtrack_iter = int(max_backtrack_iter)
    o.warm_start = int(bool(warm_start))
    o.decay_rate = float(decay_rate)
    o.nesterov = int(bool(nesterov))
    a, b = nesterov_ratio
    o.nesterov_a, o.nesterov_b = float(a), float(b)
    o.deprecated = int(bool(deprecated))
    o.dual_solver = DUAL_SOLVERS[dual_solver]
    o.trace_capacity = int(trace_capacity)
    return o


@dataclass
class BatchResult:
    """Struct-of-arrays result of a batched solve (one row per start)."""

    x: np.ndarray          # (n_starts, n_features)
    fun: np.ndarray        # (n_starts, n_objectives)
    nit: np.ndarray        # (n_starts,) int64
    status: np.ndarray     # (n_starts,) 1 converged / 0 max_iter / -1 backtracking failed
    lr: np.ndarray         # final step sizes
    nfev: np.ndarray       # objective evaluations on device
    n_dual: np.ndarray     # dual-function evaluations
    err: np.ndarray        # last max|x^k - y^k|
    time: float            # wall time of the call, seconds
    allerrs: np.ndarray | None = None
    allfuns: np.ndarray | None = None
    allvecs: np.ndarray | None = None

    @property
    def success(self) -> np.ndarray:
        return self.status == 1

    def __len__(self) -> int:
        return len(self.nit)

    def to_results(self) -> list[OptimizeResult]:
        """One ``OptimizeResult`` per start with the reference's fields."""
        out = []
        for i in range(len(self)):
            out.append(_one_result(self, i))
        return out


def _message(status: int) -> str:
    if status == 1:
        return "Optimization terminated successfully"
    if status == 0:
        return "Maximum number of iterations reached"
    return "Error: Backtracking failed to find a suitable stepsize."


def _one_result(br: BatchResult, i: int) -> OptimizeResult:
    status = int(br.status[i])
    nit = int(br.nit[i])
    fun = br.fun[i].copy()
    res = OptimizeResult(
        x=br.x[i].copy(), fun=fun if fun.shape[0] > 1 else fun[0], nit=nit,
        success=status == 1, status=status, message=_message(status),
        nfev=int(br.nfev[i]), lr=float(br.lr[i]), time=br.time,
        allvecs=None, allfuns=None, allerrs=None)
    if br.allerrs is not None:
        res.allerrs = list(br.allerrs[i, :nit])
        res.allfuns = [row if row.shape[0] > 1 else row[0] for row in br.allfuns[i, :nit + 1]]
        res.allvecs = list(br.allvecs[i, :nit + 1])
    return res


def minimize_proximal_gradient_batched(problem: Problem, X0, lr=1, tol=1e-5,
                                       tol_internal=1e-12, max_iter=1000000,
                                       max_iter_internal=100000, max_backtrack_iter=100,
                                       warm_start=False, decay_rate=0.5, nesterov=False,
                                       nesterov_ratio=(0, 0.25), return_all=False,
                                       deprecated=False, dual_solver="reference",
                                       trace_capacity=None) -> BatchResult:
    """Solve from every row of ``X0`` in ONE kernel launch (one warp per start).

    This replaces the joblib fan-out of benchmarks/benchmark.py:320-372 and of the
    PGM_experiment_with_various_a_b notebook.  ``nesterov_ratio`` may be one (a, b)
    pair or an array of shape (n_starts, 2) giving each start its own pair, which is
    how the (a, b) sweep grid is batched.  Host arrays in, host arrays out (H2D and
    D2H copies are inside the call).
    """
    if not isinstance(problem, Problem):
        raise TypeError("problem must be a zfista_b200.problems.Problem")
    X0 = _as_f64(X0)
    if X0.ndim != 2 or X0.shape[1] != problem.n_features:
        raise ValueError(f"X0 must have shape (n_starts, {problem.n_features})")
    n_starts, n = X0.shape
    m = problem.n_objectives
    ab = np.asarray(nesterov_ratio, dtype=np.float64)
    ab_arr = None
    if ab.ndim == 2:
        if ab.shape != (n_starts, 2):
            raise ValueError("per-start nesterov_ratio must have shape (n_starts, 2)")
        ab_arr = _as_f64(ab)
        pair = (0.0, 0.25)
    else:
        pair = (float(ab[0]), float(ab[1]))
    cap = 0
    if return_all:
        cap = int(trace_capacity) if trace_capacity is not None else min(int(max_iter),
                                                                       _DEFAULT_TRACE)
    t0 = time.time()
    while True:
        opts = _make_options(lr, tol, tol_internal, max_iter, max_iter_internal,
                             max_backtrack_iter, warm_start, decay_rate, nesterov, pair,
                             deprecated, dual_solver, cap)
        desc, keep = problem.descriptor()
        out = BatchResult(
            x=np.empty((n_starts, n)), fun=np.empty((n_starts, m)),
            nit=np.zeros(n_starts, dtype=np.int64), status=np.zeros(n_starts, dtype=np.int32),
            lr=np.empty(n_starts), nfev=np.zeros(n_starts, dtype=np.int64),
            n_dual=np.zeros(n_starts, dtype=np.int64), err=np.empty(n_starts), time=0.0)
        r = _lib.ZfResult()
        r.x, r.fun, r.nit, r.status = _ptr(out.x), _ptr(out.fun), _ptr(out.nit), _ptr(out.status)
        r.lr, r.nfev, r.n_dual, r.err = _ptr(out.lr), _ptr(out.nfev), _ptr(out.n_dual), _ptr(out.err)
        if cap > 0:
            out.allerrs = np.zeros((n_starts, cap))
            out.allfuns = np.zeros((n_starts, cap + 1, m))
            out.allvecs = np.zeros((n_starts, cap + 1, n))
            r.allerrs, r.allfuns, r.allvecs = _ptr(out.allerrs), _ptr(out.allfuns), _ptr(out.allvecs)
        L = _lib.lib()
        _lib.check(L.zf_solve_batched_host(C.byref(desc), C.byref(opts), n_starts, _ptr(X0),
                                           _ptr(ab_arr), C.byref(r)))
        del keep
        if cap > 0 and n_starts and int(out.nit.max()) > cap and trace_capacity is None:
            cap = int(out.nit.max())      # trace overflowed: rerun once with room for all
            continue
        break
    out.time = time.time() - t0
    return out


def minimize_proximal_gradient(f, g, jac_f, prox_wsum_g, x0, lr=1, tol=1e-5,
                               tol_internal=1e-12, max_iter=1000000,
                               max_iter_internal=100000, max_backtrack_iter=100,
                               warm_start=False, decay_rate=0.5, nesterov=False,
                               nesterov_ratio=(0, 0.25), return_all=False, verbose=False,
                               deprecated=False, dual_solver="reference") -> OptimizeResult:
    """Drop-in for ``zfista.minimize_proximal_gradient`` (proximal_gradient.py:311-555).

    Same parameters and the same ``OptimizeResult`` fields (``x``, ``fun``, ``success``,
    ``status``, ``message``, ``nit``, ``time``, ``allvecs``, ``allfuns``, ``allerrs``,
    plus ``x0``, ``tol``, ``tol_internal``, ``nesterov``, ``nesterov_ratio``), and in
    addition ``nfev`` (objective evaluations) and ``lr`` (final step).  ``f, g, jac_f,
    prox_wsum_g`` must be the bound methods of one :class:`~zfista_b200.problems.Problem`.

    ``dual_solver="reference"`` (default) follows the reference's inner solver
    (bounded Brent for two objectives); ``"newton"`` uses the exact simplex Newton
    solver for every m >= 2.
    """
    from .lasso import DenseLasso

    owner = getattr(f, "__self__", None)
    if isinstance(owner, DenseLasso):
        # the large-n path: same four closures, bound to a DenseLasso
        for name, fn in (("g", g), ("jac_f", jac_f), ("prox_wsum_g", prox_wsum_g)):
            if getattr(fn, "__self__", None) is not owner or fn.__name__ != name:
                raise TypeError("f, g, jac_f and prox_wsum_g must belong to the same DenseLasso")
        return owner.minimize_proximal_gradient(
            x0, lr=lr, tol=tol, tol_internal=tol_internal, max_iter=max_iter,
            max_iter_internal=max_iter_internal, max_backtrack_iter=max_backtrack_iter,
            warm_start=warm_start, decay_rate=decay_rate, nesterov=nesterov,
            nesterov_ratio=nesterov_ratio, return_all=return_all, verbose=verbose,
            deprecated=deprecated)
    if deprecated:
        warn("Using the deprecated option is not mathematically proven to converge. "
             "Please consider using the recommended condition instead.", stacklevel=2)
    problem = _resolve_problem(f, g, jac_f, prox_wsum_g)
    x0 = np.asarray(x0, dtype=np.float64)
    if x0.ndim != 1 or x0.shape[0] != problem.n_features:
        raise ValueError(f"len(x) should be equal to n_features, got {x0}.")
    start = time.time()
    br = minimize_proximal_gradient_batched(
        problem, x0[None, :], lr=lr, tol=tol, tol_internal=tol_internal, max_iter=max_iter,
        max_iter_internal=max_iter_internal, max_backtrack_iter=max_backtrack_iter,
        warm_start=warm_start, decay_rate=decay_rate, nesterov=nesterov,
        nesterov_ratio=nesterov_ratio, return_all=return_all or verbose,
        deprecated=deprecated, dual_solver=dual_solver)
    one = _one_result(br, 0)
    if verbose:
        print(f"|{'niter':^7}|{'max(abs(xk - yk)))':^20}|{'learning rate':^13}|")
        for k, e in enumerate(one.allerrs, start=1):
            print(f"|{k:^7}|{e:^+20.4e}|{one.lr:^13.2e}|")
    if not return_all:
        one.allvecs = one.allfuns = one.allerrs = None
    one.time = time.time() - start
    if one.status == -1:
        # proximal_gradient.py:493-509: the reference prints and returns a reduced result
        print("An error occurred: Backtracking failed to find a suitable stepsize.")
        return one
    if one.status == 0:
        warn(one.message, stacklevel=2)
    one.update(x0=x0, tol=tol, tol_internal=tol_internal, nesterov=nesterov,
               nesterov_ratio=nesterov_ratio)
    return one


def solve_subproblems(problem: Problem, Y, X_old, lr, deprecated=None, tol_internal=1e-12,
                      max_iter_internal=100000, dual_solver="reference"):
    """Batched ``_solve_subproblem`` (proximal_gradient.py:35-209): one proximal
    subproblem per row of ``Y`` / ``X_old`` with step ``lr[i]``.  Returns
    ``(x, fun, weight)`` with shapes (n, n_features), (n,), (n, n_objectives)."""
    Y, X_old = _as_f64(Y), _as_f64(X_old)
    n_items, n = Y.shape
    if X_old.shape != Y.shape or n != problem.n_features:
        raise ValueError("Y and X_old must both have shape (n, n_features)")
    lr = _as_f64(np.broadcast_to(np.asarray(lr, dtype=np.float64), (n_items,)))
    dep = None
    if deprecated is not None:
        dep = np.ascontiguousarray(np.broadcast_to(np.asarray(deprecated), (n_items,)),
                                   dtype=np.int32)
    opts = _make_options(1.0, 1e-5, tol_internal, 1, max_iter_internal, 1, False, 0.5, False,
                         (0, 0.25), False, dual_solver, 0)
    desc, keep = problem.descriptor()
    x = np.empty((n_items, n))
    fun = np.empty(n_items)
    w = np.empty((n_items, problem.n_objectives))
    _lib.check(_lib.lib().zf_solve_subproblem_host(
        C.byref(desc), C.byref(opts), n_items, _ptr(Y), _ptr(X_old), _ptr(lr), _ptr(dep),
        _ptr(x), _ptr(fun), _ptr(w)))
    del keep
    return x, fun, w
