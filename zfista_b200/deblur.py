"""Cameraman-style L1-regularised deblurring on the GPU (BASELINE configs[1]).

The reference notebook (examples/cameraman.ipynb) hands ``minimize_proximal_gradient`` four
numpy closures built from ``scipy.signal.correlate2d`` and PyWavelets' Haar transform, and
runs it once per (a, b) momentum pair under joblib.  :class:`HaarDeblurL1` is that closure
set as a device object (``csrc/zf_deblur.cu``):

    f(x) = ||R W x - b||^2,  g(x) = l1_ratio*||x||_1,  jac_f(x) = 2 W^T R (R W x - b)

with ``R = correlate2d(., kernel, mode="same", boundary="symm")`` and ``W`` the inverse
single-level 2-D Haar transform of ``x = [cA, cH, cV, cD].flatten()``.  All (a, b) pairs are
solved by ONE batched call.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import time
from warnings import warn

import numpy as np
from scipy.optimize import OptimizeResult

from . import _lib
from .problems import _as_f64, _ptr


def gaussian_kernel(size: int = 9, sigma: float = 4.0) -> np.ndarray:
    """Radial Gaussian blur kernel (what the notebook's ``skimage.filters.window(("gaussian",
    4), (9, 9))`` approximates); not normalised."""
    c = (size - 1) / 2
    i = np.arange(size) - c
    return np.exp(-(i[:, None] ** 2 + i[None, :] ** 2) / (2 * sigma ** 2))


def lipschitz_constant(kernel) -> float:
    """The notebook's ``L = 2 * max|dctn(K) / dctn(unit)|**2`` (Hansen et al. 2006): the
    Lipschitz constant of grad f for the symmetric-boundary blur; the fixed step is 1 / L."""
    from scipy.fftpack import dctn

    kernel = np.asarray(kernel, dtype=np.float64)
    unit = np.zeros(kernel.shape)
    unit[0, 0] = 1
    spectrum = dctn(kernel) / dctn(unit)
    return float(2 * np.max(np.abs(spectrum)) ** 2)


class HaarDeblurL1:
    n_objectives = 1

    def __init__(self, observed, kernel, l1_ratio: float, max_runs: int = 16):
        self.observed = _as_f64(observed)
        self.kernel = _as_f64(kernel)
        if self.observed.ndim != 2 or self.kernel.ndim != 2 or \
                self.kernel.shape[0] != self.kernel.shape[1]:
            raise ValueError("observed must be 2-d and kernel square")
        self.shape = self.observed.shape
        self.n_features = int(self.observed.size)
        self.l1_ratio = float(l1_ratio)
        self.max_runs = int(max_runs)
        self._h = C.c_void_p()
        _lib.check(_lib.lib().zf_deblur_create(
            C.byref(self._h), self.shape[0], self.shape[1], _ptr(self.kernel),
            self.kernel.shape[0], _ptr(self.observed), self.l1_ratio, self.max_runs, None))

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                _lib.lib().zf_deblur_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    # ------------------------------------------------------------------ transforms (host helpers)
    def dwt_array(self, image):
        """pywt.dwt2(image, "haar") flattened as the notebook's dwt_array (x0 construction)."""
        p = _as_f64(image)
        p00, p01, p10, p11 = p[0::2, 0::2], p[0::2, 1::2], p[1::2, 0::2], p[1::2, 1::2]
        return np.array([(p00 + p01 + p10 + p11) / 2, (p00 + p01 - p10 - p11) / 2,
                         (p00 - p01 + p10 - p11) / 2, (p00 - p01 - p10 + p11) / 2]).flatten()

    def idwt_array(self, x):
        h, w = self.shape
        cA, cH, cV, cD = _as_f64(x).reshape(4, h // 2, w // 2)
        out = np.empty((h, w))
        out[0::2, 0::2] = (cA + cH + cV + cD) / 2
        out[0::2, 1::2] = (cA + cH - cV - cD) / 2
        out[1::2, 0::2] = (cA - cH + cV - cD) / 2
        out[1::2, 1::2] = (cA - cH - cV + cD) / 2
        return out

    # ------------------------------------------------------------------ closures
    def _eval(self, x, want_jac=False):
        x = _as_f64(x)
        if x.shape != (self.n_features,):
            raise ValueError(f"len(x) should be equal to n_features, got {x}.")
        f, g = np.empty(1), np.empty(1)
        jac = np.empty(self.n_features) if want_jac else None
        _lib.check(_lib.lib().zf_deblur_eval_host(self._h, 1, _ptr(x), _ptr(f), _ptr(g), _ptr(jac)))
        return f, g, jac

    def f(self, x):
        return self._eval(x)[0]

    def g(self, x):
        return self._eval(x)[1]

    def jac_f(self, x):
        return self._eval(x, want_jac=True)[2].reshape(1, -1)

    def prox_wsum_g(self, weight, x):
        raise NotImplementedError(
            "the prox is fused into the device solve; call minimize_proximal_gradient")

    # ------------------------------------------------------------------ solver
    def minimize_proximal_gradient_batched(self, x0, nesterov_ratios, lr=1, tol=1e-5,
                                           tol_internal=1e-12, max_iter=1000000,
                                           max_backtrack_iter=100, decay_rate=0.5,
                                           nesterov=True, return_all=False, deprecated=False,
                                           trace_capacity=None):
        """One run per row of ``nesterov_ratios`` (shape (n_runs, 2)); ``x0`` is one vector
        shared by all runs or (n_runs, n).  Returns a list of OptimizeResult."""
        from .proximal_gradient import _make_options, _message

        ab = _as_f64(np.asarray(nesterov_ratios, dtype=np.float64).reshape(-1, 2))
        n_runs = len(ab)
        if n_runs > self.max_runs:
            raise ValueError(f"{n_runs} runs > max_runs={self.max_runs}")
        x0 = _as_f64(x0)
        batched = x0.ndim == 2
        if (batched and x0.shape != (n_runs, self.n_features)) or \
                (not batched and x0.shape != (self.n_features,)):
            raise ValueError("x0 must have shape (n_features,) or (n_runs, n_features)")
        t0 = time.time()
        cap = 0
        if return_all:
            cap = int(trace_capacity) if trace_capacity is not None else int(min(max_iter, 1 << 16))
        while True:
            opts = _make_options(lr, tol, tol_internal, max_iter, 100000, max_backtrack_iter,
                                 False, decay_rate, nesterov, (0, 0.25), deprecated, "reference", cap)
            n = self.n_features
            x, fun = np.empty((n_runs, n)), np.empty(n_runs)
            nit, status = np.zeros(n_runs, dtype=np.int64), np.zeros(n_runs, dtype=np.int32)
            lrs, err = np.empty(n_runs), np.empty(n_runs)
            allerrs = np.zeros((n_runs, cap)) if cap else None
            allfuns = np.zeros((n_runs, cap + 1)) if cap else None
            r = _lib.ZfResult()
            r.x, r.fun, r.nit, r.status, r.lr, r.err = (_ptr(x), _ptr(fun), _ptr(nit),
                                                        _ptr(status), _ptr(lrs), _ptr(err))
            r.allerrs, r.allfuns = _ptr(allerrs), _ptr(allfuns)
            _lib.check(_lib.lib().zf_deblur_solve_host(self._h, C.byref(opts), n_runs, _ptr(x0),
                                                       int(batched), _ptr(ab), C.byref(r)))
            if cap and int(nit.max()) > cap and trace_capacity is None:
                cap = int(nit.max())
                continue
            break
        elapsed = time.time() - t0
        out = []
        for i in range(n_runs):
            st, k = int(status[i]), int(nit[i])
            res = OptimizeResult(x=x[i], fun=np.array([fun[i]]), nit=k, success=st == 1,
                                 status=st, message=_message(st), time=elapsed, lr=float(lrs[i]),
                                 nesterov_ratio=(float(ab[i, 0]), float(ab[i, 1])),
                                 allvecs=None, allfuns=None, allerrs=None)
            if return_all:
                res.allerrs = list(allerrs[i, :k])
                res.allfuns = [np.array([v]) for v in allfuns[i, :k + 1]]
            out.append(res)
        return out

    def minimize_proximal_gradient(self, x0, nesterov=False, nesterov_ratio=(0, 0.25), **kwargs):
        """Single run with the reference's signature (proximal_gradient.py:311-331)."""
        if kwargs.get("deprecated"):
            warn("Using the deprecated option is not mathematically proven to converge. "
                 "Please consider using the recommended condition instead.", stacklevel=2)
        kwargs.pop("verbose", None)
        kwargs.pop("warm_start", None)
        kwargs.pop("max_iter_internal", None)
        res = self.minimize_proximal_gradient_batched(x0, [nesterov_ratio], nesterov=nesterov,
                                                      **kwargs)[0]
        if res.status == -1:
            print("An error occurred: Backtracking failed to find a suitable stepsize.")
        elif res.status == 0:
            warn(res.message, stacklevel=2)
        return res
