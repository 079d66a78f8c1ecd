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
# return_all: most bytes of iterates (allvecs) one launch may hold on the device, and the most the
# host may be asked to keep before the call refuses (return_all="funs" keeps F and the errors only)
_TRACE_DEVICE_BYTES = 1 << 30
_TRACE_HOST_BYTES = 16 << 30


class RaggedTrace:
    """Per-start traces of different lengths stored flat: ``t[i]`` is the array of start i
    (``(len_i,)`` or ``(len_i, width)``), ``t[i, :k]`` its first k entries."""

    def __init__(self, flat: np.ndarray, offsets: np.ndarray):
        self.flat, self.offsets = flat, offsets

    def __len__(self) -> int:
        return len(self.offsets) - 1

    def __getitem__(self, key):
        if isinstance(key, tuple):
            i, rest = key[0], key[1:]
            return self.flat[self.offsets[i]:self.offsets[i + 1]][rest if len(rest) > 1 else rest[0]]
        return self.flat[self.offsets[key]:self.offsets[key + 1]]

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    @property
    def nbytes(self) -> int:
        return int(self.flat.nbytes)


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
    trace_truncated: bool = False   # trace_capacity cut at least one start's trace short
    allerrs: RaggedTrace | None = None      # start i: (nit_i,)
    allfuns: RaggedTrace | None = None      # start i: (nit_i + 1, n_objectives), entry 0 = F(x0)
    allvecs: RaggedTrace | None = None      # start i: (nit_i + 1, n_features); None with "funs"

    @property
    def success(self) -> np.ndarray:
        return self.status == 1

    def __len__(self) -> int:
        return len(self.nit)

    def to_results(self, scalar_fun: bool = False) -> list[OptimizeResult]:
        """One ``OptimizeResult`` per start with the reference's fields."""
        return [_one_result(self, i, scalar_fun) for i in range(len(self))]


def _message(status: int) -> str:
    if status == 1:
        return "Optimization terminated successfully"
    if status == 0:
        return "Maximum number of iterations reached"
    return "Error: Backtracking failed to find a suitable stepsize."


def _one_result(br: BatchResult, i: int, scalar_fun: bool = False) -> OptimizeResult:
    """``scalar_fun``: the objective closures return scalars (LeastSquaresL1 with one objective,
    the closures of tests/test_proximal_gradient.py:49-63), so ``fun`` / ``allfuns`` are
    scalars too; every zfista.problems class returns shape-(m,) arrays, also for m = 1."""
    status = int(br.status[i])
    nit = int(br.nit[i])
    fun = br.fun[i].copy()
    res = OptimizeResult(
        x=br.x[i].copy(), fun=fun[0] if scalar_fun else fun, nit=nit,
        success=status == 1, status=status, message=_message(status),
        nfev=int(br.nfev[i]), lr=float(br.lr[i]), time=br.time,
        allvecs=None, allfuns=None, allerrs=None)
    if br.allerrs is not None:
        res.allerrs = list(br.allerrs[i])
        res.allfuns = [row[0] if scalar_fun else row for row in br.allfuns[i]]
        res.allvecs = None if br.allvecs is None else list(br.allvecs[i])
    return res


def _solve_host(problem, X0, ab_arr, opts, offsets=None, want_vecs=False):
    """One call of zf_solve_batched_host -> (BatchResult without traces, allerrs, allfuns,
    allvecs flat arrays or None)."""
    n_starts, n = X0.shape
    m = problem.n_objectives
    desc, keep = problem.descriptor()
    out = BatchResult(
        x=np.empty((n_starts, n)), fun=np.empty((n_starts, m)),
        nit=np.zeros(n_starts, dtype=np.int64), status=np.zeros(n_starts, dtype=np.int32),
        lr=np.empty(n_starts), nfev=np.zeros(n_starts, dtype=np.int64),
        n_dual=np.zeros(n_starts, dtype=np.int64), err=np.empty(n_starts), time=0.0)
    r = _lib.ZfResult()
    r.x, r.fun, r.nit, r.status = _ptr(out.x), _ptr(out.fun), _ptr(out.nit), _ptr(out.status)
    r.lr, r.nfev, r.n_dual, r.err = _ptr(out.lr), _ptr(out.nfev), _ptr(out.n_dual), _ptr(out.err)
    errs = funs = vecs = None
    if offsets is not None:
        total = int(offsets[-1])
        errs = np.zeros(total)
        funs = np.zeros((total + n_starts, m))
        r.allerrs, r.allfuns, r.trace_offsets = _ptr(errs), _ptr(funs), _ptr(offsets)
        if want_vecs:
            vecs = np.zeros((total + n_starts, n))
            r.allvecs = _ptr(vecs)
    _lib.check(_lib.lib().zf_solve_batched_host(C.byref(desc), C.byref(opts), n_starts, _ptr(X0),
                                                _ptr(ab_arr), C.byref(r)))
    del keep
    return out, errs, funs, vecs


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

    ``return_all`` (benchmark.py passes it for every start): ``True`` records ``allerrs``,
    ``allfuns`` and ``allvecs`` as the reference does, ``"funs"`` leaves the iterates out.  The
    traces are RAGGED (:class:`RaggedTrace`): a first launch without traces gives every start's
    iteration count, a second one (the solver is deterministic) writes each start's trace into
    exactly the space it needs -- nothing of size ``n_starts x capacity x n_features`` is ever
    allocated.  The iterates go through the device in groups of starts of at most 1 GiB.
    ``trace_capacity`` truncates every start's trace to its first ``trace_capacity`` iterations
    (``BatchResult.trace_truncated`` says whether that cut anything).
    """
    if not isinstance(problem, Problem):
        raise TypeError("problem must be a zfista_b200.problems.Problem")
    X0 = _as_f64(X0)
    if X0.ndim != 2 or X0.shape[1] != problem.n_features:
        raise ValueError(f"X0 must have shape (n_starts, {problem.n_features})")
    if return_all not in (False, True, "funs"):
        raise ValueError('return_all must be False, True or "funs"')
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
    t0 = time.time()
    opts = _make_options(lr, tol, tol_internal, max_iter, max_iter_internal, max_backtrack_iter,
                         warm_start, decay_rate, nesterov, pair, deprecated, dual_solver, 0)
    out, _, _, _ = _solve_host(problem, X0, ab_arr, opts)
    if return_all and n_starts:
        want_vecs = return_all is True
        lens = out.nit.copy()
        if trace_capacity is not None:
            lens = np.minimum(lens, int(trace_capacity))
        out.trace_truncated = bool(np.any(lens < out.nit))
        offsets = np.zeros(n_starts + 1, dtype=np.int64)
        np.cumsum(lens, out=offsets[1:])
        total = int(offsets[-1])
        vec_bytes = (total + n_starts) * n * 8 if want_vecs else 0
        if vec_bytes > _TRACE_HOST_BYTES:
            raise MemoryError(
                f"return_all=True would keep {vec_bytes / 2**30:.1f} GiB of iterates (allvecs) on "
                'the host; pass return_all="funs" for allfuns / allerrs only, or trace_capacity')
        errs = np.zeros(total)
        funs = np.zeros((total + n_starts, m))
        vecs = np.zeros((total + n_starts, n)) if want_vecs else None
        # groups of consecutive starts whose iterates fit the device budget
        per_start = (lens + 1) * n * 8 if want_vecs else np.zeros(n_starts, dtype=np.int64)
        lo = 0
        while lo < n_starts:
            hi, acc = lo, 0
            while hi < n_starts and (hi == lo or acc + per_start[hi] <= _TRACE_DEVICE_BYTES):
                acc += int(per_start[hi])
                hi += 1
            goff = np.ascontiguousarray(offsets[lo:hi + 1] - offsets[lo])
            part, e, f_, v = _solve_host(problem, X0[lo:hi],
                                         None if ab_arr is None else ab_arr[lo:hi], opts, goff,
                                         want_vecs)
            if not np.array_equal(part.nit, out.nit[lo:hi]):
                raise RuntimeError("the trace pass did not reproduce the first pass's iterations")
            errs[offsets[lo]:offsets[hi]] = e
            funs[offsets[lo] + lo:offsets[hi] + hi] = f_
            if want_vecs:
                vecs[offsets[lo] + lo:offsets[hi] + hi] = v
            lo = hi
        out.allerrs = RaggedTrace(errs, offsets)
        foff = offsets + np.arange(n_starts + 1)
        out.allfuns = RaggedTrace(funs, foff)
        out.allvecs = RaggedTrace(vecs, foff) if want_vecs else None
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
        nesterov_ratio=nesterov_ratio,
        return_all=return_all if return_all else ("funs" if verbose else False),
        deprecated=deprecated, dual_solver=dual_solver)
    from .problems import LeastSquaresL1

    one = _one_result(br, 0, scalar_fun=isinstance(problem, LeastSquaresL1)
                      and problem.n_objectives == 1)
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
