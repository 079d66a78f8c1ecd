"""Problem classes with the constructor signatures of ``zfista.problems``
(/root/reference/zfista/problems.py), backed by device functors.

A ``Problem`` here is a *descriptor*: ``f``, ``jac_f``, ``g`` and ``prox_wsum_g`` are
evaluated by the CUDA functors in ``csrc/zf_problems.cuh`` through the C ABI
(``zf_problem_eval_host``), and ``minimize_proximal_gradient`` runs the whole
FISTA/ISTA loop on the GPU (``zf_solve_batched_host``).  Nothing is computed on the
CPU; without the CUDA library or a GPU the calls raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Sequence

import numpy as np

from . import _lib

KIND_IDS = {
    "JOS1": 0, "SD": 1, "FDS": 2, "ZDT1": 3, "TOI4": 4, "TRIDIA": 5,
    "LinearFunctionRank1": 6, "LeastSquaresL1": 7,
}


def _as_f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Problem:
    """Mirror of ``zfista.problems.Problem`` (problems.py:25-150).

    Parameters are those of the reference: ``n_features``, ``n_objectives``,
    ``l1_ratios``, ``l1_shifts`` (one per objective) and ``bounds`` = (lower, upper),
    each a scalar or an array of shape (n_features,).
    """

    def __init__(self, n_features: int, n_objectives: int,
                 l1_ratios: Sequence[float] | None = None,
                 l1_shifts: Sequence[float] | None = None,
                 bounds: tuple | None = None) -> None:
        self.n_features = int(n_features)
        self.n_objectives = int(n_objectives)
        self.l1_ratios = None if l1_ratios is None else np.array(l1_ratios)
        self.l1_shifts = (np.zeros(n_objectives) if l1_shifts is None
                          else np.array(l1_shifts))
        self.bounds = bounds
        self.name = self._generate_name()

    # problems.py:81-91
    def _generate_name(self) -> str:
        parts = [type(self).__name__, f"n_{self.n_features}"]
        if self.l1_ratios is not None:
            parts.append("l1_ratios_" + "_".join(map(str, self.l1_ratios)))
            parts.append("l1_shifts_" + "_".join(map(str, self.l1_shifts)))
        if self.bounds is not None:
            parts.append("bounds_" + "_".join(map(str, [self.bounds[0], self.bounds[1]])))
        return "_".join(parts)

    # ------------------------------------------------------------------ descriptor
    def _kind(self) -> int:
        return KIND_IDS[type(self).__name__]

    def descriptor(self):
        """(ZfProblem, keepalive): the C-ABI descriptor with HOST pointers."""
        if self.n_objectives > _lib.ZF_MAX_OBJECTIVES:
            raise NotImplementedError(
                f"n_objectives={self.n_objectives}: the device dual solver supports up to "
                f"{_lib.ZF_MAX_OBJECTIVES} objectives")
        p = _lib.ZfProblem()
        keep: list[Any] = []
        p.kind = self._kind()
        p.n_features = self.n_features
        p.n_objectives = self.n_objectives
        p.has_l1 = int(self.l1_ratios is not None)
        if self.l1_ratios is not None:
            # problems.py:108-111
            if self.n_objectives != len(self.l1_ratios):
                raise ValueError("len(l1_ratios) should be equal to n_objectives.")
            if self.n_objectives != len(self.l1_shifts):
                raise ValueError("len(l1_shifts) should be equal to n_objectives.")
            for i in range(self.n_objectives):
                p.l1_ratios[i] = float(self.l1_ratios[i])
                p.l1_shifts[i] = float(self.l1_shifts[i])
        p.has_bounds = int(self.bounds is not None)
        if self.bounds is not None:
            lo, hi = self.bounds
            if np.ndim(lo) == 0 and np.ndim(hi) == 0:
                p.bounds_are_arrays = 0
                p.lower, p.upper = float(lo), float(hi)
            else:
                lo_v = _as_f64(np.broadcast_to(np.asarray(lo, dtype=np.float64),
                                               (self.n_features,)))
                hi_v = _as_f64(np.broadcast_to(np.asarray(hi, dtype=np.float64),
                                               (self.n_features,)))
                keep += [lo_v, hi_v]
                p.bounds_are_arrays = 1
                p.lower_v, p.upper_v = _ptr(lo_v), _ptr(hi_v)
        self._fill_extra(p, keep)
        return p, keep

    def _fill_extra(self, p, keep) -> None:
        pass

    # ------------------------------------------------------------------ functors
    def _check_x(self, x) -> np.ndarray:
        if self.n_features != len(x):
            raise ValueError(f"len(x) should be equal to n_features, got {x}.")
        return _as_f64(x)

    def _eval(self, x, weight=None, want=("f",)):
        x = self._check_x(x)
        p, keep = self.descriptor()
        m, n = self.n_objectives, self.n_features
        out = {k: None for k in ("f", "g", "jac", "prox")}
        if "f" in want:
            out["f"] = np.empty(m)
        if "g" in want:
            out["g"] = np.empty(m)
        if "jac" in want:
            out["jac"] = np.empty((m, n))
        w = None
        if "prox" in want:
            w = _as_f64(weight)
            out["prox"] = np.empty(n)
        L = _lib.lib()
        _lib.check(L.zf_problem_eval_host(C.byref(p), 1, _ptr(x), _ptr(w), _ptr(out["f"]),
                                          _ptr(out["g"]), _ptr(out["jac"]), _ptr(out["prox"])))
        return out

    def f(self, x):
        return self._eval(x, want=("f",))["f"]

    def jac_f(self, x):
        return self._eval(x, want=("jac",))["jac"]

    def g(self, x):
        return self._eval(x, want=("g",))["g"]

    def prox_wsum_g(self, weight, x):
        if self.n_features != len(x):
            raise ValueError(f"len(x) should be equal to n_features, got {x}.")
        if self.n_objectives != len(weight):
            raise ValueError("len(weight) should be equal to n_objectives.")
        return self._eval(x, weight=weight, want=("prox",))["prox"]

    # ------------------------------------------------------------------ solvers
    def minimize_proximal_gradient(self, x0, **kwargs: Any):
        """problems.py:140-150: one start, same keyword arguments as the function."""
        from .proximal_gradient import minimize_proximal_gradient

        return minimize_proximal_gradient(self.f, self.g, self.jac_f, self.prox_wsum_g, x0,
                                          **kwargs)

    def minimize_proximal_gradient_batched(self, X0, **kwargs: Any):
        """All rows of X0 (n_starts, n_features) in one kernel launch; see
        :func:`zfista_b200.proximal_gradient.minimize_proximal_gradient_batched`."""
        from .proximal_gradient import minimize_proximal_gradient_batched

        return minimize_proximal_gradient_batched(self, X0, **kwargs)


class JOS1(Problem):
    """f1 = ||x||^2 / n, f2 = ||x - 2||^2 / n   (problems.py:153-205)."""

    def __init__(self, n_features: int = 5, l1_ratios=None, l1_shifts=None, bounds=None):
        super().__init__(n_features=n_features, n_objectives=2, l1_ratios=l1_ratios,
                         l1_shifts=l1_shifts, bounds=bounds)


class SD(Problem):
    """Stadler-Dauer four-bar truss, n = 4, box (1e-6, inf)   (problems.py:208-264)."""

    def __init__(self) -> None:
        super().__init__(n_features=4, n_objectives=2, bounds=(1e-6, np.inf))


class FDS(Problem):
    """Fliege-Drummond-Svaiter tri-objective   (problems.py:267-328)."""

    def __init__(self, n_features: int = 10, l1_ratios=None, l1_shifts=None, bounds=None):
        super().__init__(n_features=n_features, n_objectives=3, l1_ratios=l1_ratios,
                         l1_shifts=l1_shifts, bounds=bounds)


class ZDT1(Problem):
    """ZDT1 with box (1e-6, inf)   (problems.py:331-386)."""

    def __init__(self, n_features: int = 30) -> None:
        super().__init__(n_features=n_features, n_objectives=2, bounds=(1e-6, np.inf))


class TOI4(Problem):
    """Toint problem 4, n = 4   (problems.py:389-448)."""

    def __init__(self, l1_ratios=None, l1_shifts=None, bounds=None):
        super().__init__(n_features=4, n_objectives=2, l1_ratios=l1_ratios,
                         l1_shifts=l1_shifts, bounds=bounds)


class TRIDIA(Problem):
    """Toint tridiagonal, n = 3, m = 3   (problems.py:451-514)."""

    def __init__(self, l1_ratios=None, l1_shifts=None, bounds=None):
        super().__init__(n_features=3, n_objectives=3, l1_ratios=l1_ratios,
                         l1_shifts=l1_shifts, bounds=bounds)


class LinearFunctionRank1(Problem):
    """f_i = (i * sum_j j x_j - 1)^2   (problems.py:517-578)."""

    def __init__(self, n_features: int = 10, n_objectives: int = 4, l1_ratios=None,
                 l1_shifts=None, bounds=None):
        super().__init__(n_features=n_features, n_objectives=n_objectives,
                         l1_ratios=l1_ratios, l1_shifts=l1_shifts, bounds=bounds)


class LeastSquaresL1(Problem):
    """``scale * ||A x - b||^2 + l1_ratio * ||x||_1`` as a device functor.

    This is the closure set every single-objective example of the reference writes by
    hand (tests/test_proximal_gradient.py:49-63 with scale = 1/6,
    examples/cameraman.ipynb with scale = 1).  ``n_objectives`` > 1 replicates the same
    objective, as the bi-/tri-objective toy tests do (test_proximal_gradient.py:113-213).
    For large dense A use :class:`zfista_b200.lasso.DenseLasso`, which streams A from
    HBM instead of solving inside one warp.
    """

    def __init__(self, A, b, l1_ratio: float, scale: float = 1.0, n_objectives: int = 1):
        self.A = _as_f64(np.atleast_2d(A))
        self.b = _as_f64(b)
        if self.A.shape[0] != self.b.shape[0]:
            raise ValueError("A and b have inconsistent shapes")
        self.l1_ratio = float(l1_ratio)
        self.scale = float(scale)
        super().__init__(n_features=self.A.shape[1], n_objectives=n_objectives)

    def _fill_extra(self, p, keep) -> None:
        keep += [self.A, self.b]
        p.A, p.b = _ptr(self.A), _ptr(self.b)
        p.n_rows = self.A.shape[0]
        p.scale = self.scale
        p.l1 = self.l1_ratio

    # single-objective closures return scalars / flat gradients like the reference tests
    def f(self, x):
        v = super().f(x)
        return v[0] if self.n_objectives == 1 else v

    def g(self, x):
        v = super().g(x)
        return v[0] if self.n_objectives == 1 else v

    def jac_f(self, x):
        j = super().jac_f(x)
        return j[0] if self.n_objectives == 1 else j

    def prox_wsum_g(self, weight, x):
        w = np.atleast_1d(np.asarray(weight, dtype=np.float64))
        if self.n_features != len(x):
            raise ValueError(f"len(x) should be equal to n_features, got {x}.")
        if self.n_objectives != len(w):
            raise ValueError("len(weight) should be equal to n_objectives.")
        return self._eval(x, weight=w, want=("prox",))["prox"]
