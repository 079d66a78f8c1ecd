"""Large-n single-objective LASSO path: ``scale*||A x - b||^2 + l1_ratio*||x||_1`` with a
dense fp64 ``A`` resident in HBM (BASELINE north_star (c)).

The reference solves this by handing ``minimize_proximal_gradient`` four numpy closures
(tests/test_proximal_gradient.py:49-63; examples/cameraman.ipynb writes the same four for
its blur operator).  :class:`DenseLasso` is that closure set as a device object: its
``f / g / jac_f / prox_wsum_g`` are the same functions, evaluated by the CUDA kernels of
``csrc/zf_lasso.cu``, and ``minimize_proximal_gradient`` runs the FISTA / ISTA loop of
proximal_gradient.py:474-555 with every pass over ``A`` on the GPU.

PyTorch is only the buffer carrier (device allocation, streams, and -- for the row-sharded
multi-GPU form -- ``torch.distributed`` all-reduce of ``A^T r``).  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import time
from warnings import warn

import numpy as np
from scipy.optimize import OptimizeResult

from . import _lib


def _torch():
    import torch

    return torch


class _DevView:
    """Expose a raw device pointer to torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {
            "shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2, "strides": None}


class DenseLasso:
    """``f(x) = scale*||A x - b||^2``, ``g(x) = l1_ratio*||x||_1`` on the GPU.

    Parameters
    ----------
    A : (n_rows, n_cols) float64, numpy array or CUDA torch tensor (row-major).  With
        ``process_group`` set this is the calling rank's block of rows.
    b : (n_rows,) float64, same kind.
    l1_ratio, scale : the constants of the reference closures (scale = 1/(2*n_samples) in
        tests/test_proximal_gradient.py, 1 in the cameraman notebook).
    process_group : optional ``torch.distributed`` group; rows are sharded over its ranks
        and ``A^T r`` / ``||r||^2`` are summed with one all-reduce per pass.
    """

    n_objectives = 1

    def __init__(self, A, b, l1_ratio: float, scale: float = 1.0, device=None,
                 process_group=None, distributed: bool = False):
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("DenseLasso needs a CUDA device (zfista_b200 has no CPU fallback)")
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self.A = self._to_dev(A, 2)
        self.b = self._to_dev(b, 1)
        if self.A.shape[0] != self.b.shape[0]:
            raise ValueError("A and b have inconsistent shapes")
        self.n_rows, self.n_features = int(self.A.shape[0]), int(self.A.shape[1])
        self.l1_ratio = float(l1_ratio)
        self.scale = float(scale)
        self.group = process_group
        self.distributed = bool(distributed or process_group is not None)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            self._stream = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().zf_lasso_create(
                C.byref(self._h), C.c_void_p(self.A.data_ptr()), C.c_void_p(self.b.data_ptr()),
                self.n_rows, self.n_features, self.scale, self.l1_ratio,
                C.c_void_p(self._stream)))
        n = C.c_int64()
        ptr = _lib.lib().zf_lasso_partial(self._h, C.byref(n))
        self._partial = torch.as_tensor(_DevView(int(ptr), int(n.value)), device=self.device)

    def _to_dev(self, a, ndim):
        torch = _torch()
        if isinstance(a, torch.Tensor):
            t = a.to(device=self.device, dtype=torch.float64).contiguous()
        else:
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64))).to(self.device)
        if t.dim() != ndim:
            raise ValueError(f"expected a {ndim}-d array")
        return t

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                _lib.lib().zf_lasso_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    # ------------------------------------------------------------------ closures
    def _allreduce(self):
        if self.distributed:
            import torch.distributed as dist

            dist.all_reduce(self._partial, group=self.group)

    def hbm_passes_per_gradient(self) -> int:
        """1 if this shape uses a fused one-pass A^T(Av - b) kernel, 2 for the two-pass form."""
        return int(_lib.lib().zf_lasso_passes(self._h))

    def gradient(self, x):
        """(grad f(x), f(x)) as device tensors: one residual pass + one A^T pass."""
        torch = _torch()
        xd = self._to_dev(x, 1)
        if xd.shape[0] != self.n_features:
            raise ValueError(f"len(x) should be equal to n_features, got {x}.")
        grad = torch.empty(self.n_features, dtype=torch.float64, device=self.device)
        fval = torch.empty(1, dtype=torch.float64, device=self.device)
        if self.distributed:
            raise NotImplementedError("gradient() is the single-GPU convenience entry")
        _lib.check(_lib.lib().zf_lasso_gradient_device(
            self._h, C.c_void_p(xd.data_ptr()), C.c_void_p(grad.data_ptr()),
            C.c_void_p(fval.data_ptr())))
        return grad, fval

    def f(self, x):
        return float(self.gradient(x)[1].item())

    def jac_f(self, x):
        return self.gradient(x)[0].cpu().numpy()

    def g(self, x):
        return self.l1_ratio * float(self._to_dev(x, 1).abs().sum().item())

    def prox_wsum_g(self, weight, x):
        # only used when a caller evaluates the closure directly; the solver's prox is on device
        torch = _torch()
        xd = self._to_dev(x, 1)
        t = self.l1_ratio * float(np.asarray(weight).reshape(-1)[0])
        return (torch.sign(xd) * torch.clamp(xd.abs() - t, min=0.0)).cpu().numpy()

    # ------------------------------------------------------------------ solver
    def minimize_proximal_gradient(self, x0, lr=1, tol=1e-5, tol_internal=1e-12,
                                   max_iter=1000000, max_iter_internal=100000,
                                   max_backtrack_iter=100, warm_start=False, decay_rate=0.5,
                                   nesterov=False, nesterov_ratio=(0, 0.25), return_all=False,
                                   verbose=False, deprecated=False, trace_capacity=None,
                                   return_device=False) -> OptimizeResult:
        """Same keyword arguments and OptimizeResult fields as the reference
        (proximal_gradient.py:311-555).  ``allvecs`` is not recorded on this path (n is
        large); ``allerrs`` / ``allfuns`` are."""
        from .proximal_gradient import _make_options, _message

        torch = _torch()
        if deprecated:
            warn("Using the deprecated option is not mathematically proven to converge. "
                 "Please consider using the recommended condition instead.", stacklevel=2)
        start = time.time()
        x0d = self._to_dev(x0, 1)
        if x0d.shape[0] != self.n_features:
            raise ValueError(f"len(x) should be equal to n_features, got {x0}.")
        cap = 0
        if return_all or verbose:
            cap = int(trace_capacity) if trace_capacity is not None else int(min(max_iter, 1 << 20))
        opts = _make_options(lr, tol, tol_internal, max_iter, max_iter_internal,
                             max_backtrack_iter, warm_start, decay_rate, nesterov,
                             nesterov_ratio, deprecated, "reference", cap)
        allerrs = np.zeros(cap) if cap else None
        allfuns = np.zeros(cap + 1) if cap else None
        xd = torch.empty_like(x0d)
        fun = C.c_double()
        nit = C.c_int64()
        status = C.c_int32()
        L = _lib.lib()
        with torch.cuda.device(self.device):
            if not self.distributed:
                _lib.check(L.zf_lasso_solve(
                    self._h, C.byref(opts), C.c_void_p(x0d.data_ptr()), C.c_void_p(xd.data_ptr()),
                    C.byref(fun), C.byref(nit), C.byref(status),
                    None if allerrs is None else allerrs.ctypes.data_as(C.c_void_p),
                    None if allfuns is None else allfuns.ctypes.data_as(C.c_void_p)))
            else:
                if cap:
                    raise NotImplementedError("return_all is not available on the row-sharded path")
                from .distributed import run_split_lasso

                h = self._h

                class _Ops:
                    def begin(self_):
                        _lib.check(L.zf_lasso_begin(h, C.byref(opts), C.c_void_p(x0d.data_ptr())))

                    def grad(self_, which):
                        _lib.check(L.zf_lasso_grad(h, int(which)))

                    def step(self_):
                        nxt = C.c_int32(0)
                        _lib.check(L.zf_lasso_step(h, C.byref(nxt)))
                        return nxt.value

                    def finish(self_):
                        _lib.check(L.zf_lasso_finish(h, C.c_void_p(xd.data_ptr()), C.byref(fun),
                                                     C.byref(nit), C.byref(status)))

                run_split_lasso(_Ops(), self._allreduce)
        st, k = int(status.value), int(nit.value)
        res = OptimizeResult(
            x=xd if return_device else xd.cpu().numpy(), fun=float(fun.value), nit=k,
            success=st == 1, status=st, message=_message(st), time=time.time() - start,
            allvecs=None, allfuns=None, allerrs=None)
        if return_all:
            res.allerrs = list(allerrs[:k])
            res.allfuns = list(allfuns[:k + 1])
        if verbose:
            print(f"|{'niter':^7}|{'max(abs(xk - yk)))':^20}|")
            for i, e in enumerate(allerrs[:k], start=1):
                print(f"|{i:^7}|{e:^+20.4e}|")
        if st == -1:
            print("An error occurred: Backtracking failed to find a suitable stepsize.")
            return res
        if st == 0:
            warn(res.message, stacklevel=2)
        res.update(x0=x0, tol=tol, tol_internal=tol_internal, nesterov=nesterov,
                   nesterov_ratio=nesterov_ratio)
        return res
