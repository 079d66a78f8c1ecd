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
import os
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
        self.peer_exchange = False
        if self.distributed:
            self._setup_peer_exchange()

    def _setup_peer_exchange(self):
        """One node, <= 8 ranks: map every rank's exchange buffer into every other rank
        (cudaIpc handles all-gathered over the control plane) so that the all-reduce of
        ``[A^T r | sum r^2]`` becomes peer-memory reads inside the prox kernel
        (csrc/zf_lasso.cu, P2PBuf).  Falls back to the NCCL all-reduce -- on every rank, by
        agreement -- if any rank cannot map its peers (ZF_LASSO_P2P=0 forces the fallback)."""
        import os

        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()):
            return
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        L = _lib.lib()
        buf = C.create_string_buffer(64)
        ok = os.environ.get("ZF_LASSO_P2P", "1") != "0" and 1 <= world <= 8
        with _torch().cuda.device(self.device):
            if ok:
                ok = L.zf_lasso_p2p_export(self._h, buf) == 0
            handles = [None] * world
            dist.all_gather_object(handles, (bool(ok), bytes(buf.raw)), group=self.group)
            ok = all(h[0] for h in handles)
            if ok:
                blob = b"".join(h[1] for h in handles)
                ok = L.zf_lasso_p2p_attach(self._h, rank, world, blob) == 0
            agreed = [None] * world
            dist.all_gather_object(agreed, bool(ok), group=self.group)
            if not all(agreed):
                L.zf_lasso_p2p_attach(self._h, 0, 0, None)      # world = 0: switch it off again
        self.peer_exchange = all(agreed) and bool(L.zf_lasso_p2p_active(self._h))

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
        if self.distributed and not self.peer_exchange:
            import torch.distributed as dist

            dist.all_reduce(self._partial, group=self.group)

    def _allreduce_ss(self):
        """all-reduce of the residual-norm partial alone (the last value of ``partial``)"""
        if self.distributed and not self.peer_exchange:
            import torch.distributed as dist

            dist.all_reduce(self._partial[self.n_features:], group=self.group)

    def _use_current_stream(self):
        """The library enqueues on the stream it is told; tensors handed in and the NCCL
        all-reduce are ordered on torch's CURRENT stream, so hand that one over at every entry."""
        torch = _torch()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        if stream != self._stream:
            _lib.check(_lib.lib().zf_lasso_set_stream(self._h, C.c_void_p(stream)))
            self._stream = stream

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
        with torch.cuda.device(self.device):
            self._use_current_stream()
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
                                   return_device=False, allvecs_bytes=512 << 20) -> OptimizeResult:
        """Same keyword arguments and OptimizeResult fields as the reference
        (proximal_gradient.py:311-555).  ``return_all=True`` records ``allerrs``, ``allfuns``
        and -- as long as they fit ``allvecs_bytes`` (default 512 MiB; n is large on this path)
        -- the iterates ``allvecs``; ``return_all="funs"`` leaves the iterates out."""
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
        # iterates: at most allvecs_bytes of them (rows x^0 .. x^{vec_rows - 1})
        vec_rows = 0
        if return_all is True and cap:
            vec_rows = int(min(cap + 1, max(1, allvecs_bytes // (8 * self.n_features))))
        allvecs = np.zeros((vec_rows, self.n_features)) if vec_rows else None
        xd = torch.empty_like(x0d)
        fun = C.c_double()
        nit = C.c_int64()
        status = C.c_int32()
        L = _lib.lib()
        with torch.cuda.device(self.device):
            self._use_current_stream()
            if allvecs is not None:
                _lib.check(L.zf_lasso_set_allvecs(self._h, allvecs.ctypes.data_as(C.c_void_p),
                                                  vec_rows))
            if not self.distributed:
                _lib.check(L.zf_lasso_solve(
                    self._h, C.byref(opts), C.c_void_p(x0d.data_ptr()), C.c_void_p(xd.data_ptr()),
                    C.byref(fun), C.byref(nit), C.byref(status),
                    None if allerrs is None else allerrs.ctypes.data_as(C.c_void_p),
                    None if allfuns is None else allfuns.ctypes.data_as(C.c_void_p)))
            else:
                from .distributed import run_device_lasso

                h = self._h
                want_trace = 1 if cap else 0
                pe = None if allerrs is None else allerrs.ctypes.data_as(C.c_void_p)
                pf = None if allfuns is None else allfuns.ctypes.data_as(C.c_void_p)

                class _Ops:
                    def begin(self_):
                        _lib.check(L.zf_lasso_dev_begin(h, C.byref(opts), C.c_void_p(x0d.data_ptr()),
                                                        1, want_trace))

                    def stage(self_, k):
                        _lib.check(L.zf_lasso_dev_stage(h, int(k)))

                    def needs_feval(self_):
                        return bool(L.zf_lasso_dev_needs_feval(h))

                    def snapshot(self_, slot):
                        _lib.check(L.zf_lasso_dev_poll(h, int(slot), 0, None, None))

                    def wait(self_, slot):
                        done = C.c_int32(0)
                        _lib.check(L.zf_lasso_dev_poll(h, int(slot), 1, C.byref(done), None))
                        return bool(done.value)

                    def finish(self_):
                        _lib.check(L.zf_lasso_dev_finish(h, C.c_void_p(xd.data_ptr()), C.byref(fun),
                                                         C.byref(nit), C.byref(status), None, pe, pf))

                if self.peer_exchange:
                    # the peer exchange spins on the peers' flags inside kernels (bounded, ~10 s):
                    # line the ranks up first so that a rank that was busy elsewhere is not
                    # mistaken for a dead one
                    import torch.distributed as dist

                    dist.barrier(group=self.group)
                run_device_lasso(_Ops(), self._allreduce, self._allreduce_ss)
        st, k = int(status.value), int(nit.value)
        if st == -3:
            raise RuntimeError("row-sharded LASSO: a peer rank never published its partials "
                               "(peer-memory exchange timed out)")
        res = OptimizeResult(
            x=xd if return_device else xd.cpu().numpy(), fun=float(fun.value), nit=k,
            success=st == 1, status=st, message=_message(st), time=time.time() - start,
            allvecs=None, allfuns=None, allerrs=None)
        if return_all:
            res.allerrs = list(allerrs[:k])
            res.allfuns = list(allfuns[:k + 1])
            if allvecs is not None:
                res.allvecs = list(allvecs[:min(k + 1, vec_rows)])
                res.allvecs_truncated = k + 1 > vec_rows
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


class DenseLassoMulti:
    """``n_runs`` LASSO runs that share one dense ``A``:  run k minimises
    ``scale*||A x - b_k||^2 + l1_ratio*||x||_1`` from ``x0_k`` with momentum ``(a_k, b_k)``.

    This is the reference's joblib fan-out of ``minimize_proximal_gradient`` over (a, b)
    pairs / starting points / observations (examples/PGM_experiment_with_various_a_b.ipynb
    ``run()``, examples/cameraman.ipynb) for the dense closures of
    tests/test_proximal_gradient.py:49-63.  All runs advance in lockstep and one gradient of
    every run is two FP64 tensor-core DGEMM passes over ``A`` (csrc/zf_lasso_multi.cu), so ``A``
    crosses HBM ``2/n_runs`` times per run and iteration instead of once.

    Parameters
    ----------
    A : (n_rows, n_cols) float64, n_cols even; numpy array or CUDA tensor.  With
        ``process_group`` set this is the calling rank's block of rows.
    b : (n_rows,) shared by all runs, or (n_runs, n_rows).
    n_runs : 1..32.
    """

    n_objectives = 1
    MAX_RUNS = 32

    def __init__(self, A, b, l1_ratio: float, n_runs: int, scale: float = 1.0, device=None,
                 process_group=None, distributed: bool = False):
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("DenseLassoMulti needs a CUDA device (zfista_b200 has no CPU fallback)")
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self.n_runs = int(n_runs)
        if not 1 <= self.n_runs <= self.MAX_RUNS:
            raise ValueError(f"n_runs must be in 1..{self.MAX_RUNS}")
        self.A = DenseLasso._to_dev(self, A, 2)
        bt = b if isinstance(b, torch.Tensor) else np.asarray(b, dtype=np.float64)
        self.b = DenseLasso._to_dev(self, bt, bt.ndim if not isinstance(bt, torch.Tensor) else bt.dim())
        self.n_rows, self.n_features = int(self.A.shape[0]), int(self.A.shape[1])
        self.b_batched = self.b.dim() == 2
        if self.b.shape[-1] != self.n_rows or (self.b_batched and self.b.shape[0] != self.n_runs):
            raise ValueError("b must have shape (n_rows,) or (n_runs, n_rows)")
        self.l1_ratio = float(l1_ratio)
        self.scale = float(scale)
        self.group = process_group
        self.distributed = bool(distributed or process_group is not None)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            self._stream = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().zf_lasso_multi_create(
                C.byref(self._h), C.c_void_p(self.A.data_ptr()), self.n_rows, self.n_features,
                C.c_void_p(self.b.data_ptr()), int(self.b_batched), self.n_runs, self.scale,
                self.l1_ratio, C.c_void_p(self._stream)))
        n = C.c_int64()
        ptr = _lib.lib().zf_lasso_multi_partial(self._h, C.byref(n))
        self._partial = torch.as_tensor(_DevView(int(ptr), int(n.value)), device=self.device)
        kp, pitch = C.c_int64(0), C.c_int64(0)
        _lib.check(_lib.lib().zf_lasso_multi_layout(self._h, C.byref(kp), C.byref(pitch)))
        self._ss_offset = int(kp.value) * int(pitch.value)

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                _lib.lib().zf_lasso_multi_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def _allreduce(self):
        if self.distributed:
            import torch.distributed as dist

            dist.all_reduce(self._partial, group=self.group)

    def _allreduce_ss(self):
        """all-reduce of the residual norms alone (the last kp values of ``partial``)"""
        if self.distributed:
            import torch.distributed as dist

            dist.all_reduce(self._partial[self._ss_offset:], group=self.group)

    def _use_current_stream(self):
        """as DenseLasso._use_current_stream: the library enqueues on torch's CURRENT stream"""
        stream = _torch().cuda.current_stream(self.device).cuda_stream
        if stream != self._stream:
            _lib.check(_lib.lib().zf_lasso_multi_set_stream(self._h, C.c_void_p(stream)))
            self._stream = stream

    def _x_batch(self, x):
        """(tensor, is_batched): one vector shared by all runs or (n_runs, n_features)."""
        torch = _torch()
        nd = x.dim() if isinstance(x, torch.Tensor) else np.asarray(x).ndim
        xd = DenseLasso._to_dev(self, x, nd)
        if tuple(xd.shape) not in ((self.n_features,), (self.n_runs, self.n_features)):
            raise ValueError("x must have shape (n_features,) or (n_runs, n_features)")
        return xd, xd.dim() == 2

    def gradient(self, X):
        """(grad f_k(x_k), f_k(x_k)) for every run as device tensors (n_runs, n_features) and
        (n_runs,): the two DGEMM passes."""
        torch = _torch()
        if self.distributed:
            raise NotImplementedError("gradient() is the single-GPU convenience entry")
        xd, batched = self._x_batch(X)
        if not batched:
            xd = xd.expand(self.n_runs, -1).contiguous()
        grad = torch.empty(self.n_runs, self.n_features, dtype=torch.float64, device=self.device)
        fval = torch.empty(self.n_runs, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            self._use_current_stream()
            _lib.check(_lib.lib().zf_lasso_multi_gradient_device(
                self._h, C.c_void_p(xd.data_ptr()), C.c_void_p(grad.data_ptr()),
                C.c_void_p(fval.data_ptr())))
        return grad, fval

    def minimize_proximal_gradient_batched(self, x0, nesterov_ratios=None, lr=1, tol=1e-5,
                                           tol_internal=1e-12, max_iter=1000000,
                                           max_backtrack_iter=100, decay_rate=0.5,
                                           nesterov=False, nesterov_ratio=(0, 0.25),
                                           return_all=False, deprecated=False,
                                           trace_capacity=None, return_device=False):
        """One run per row of ``nesterov_ratios`` (``(n_runs, 2)``; default: ``nesterov_ratio``
        for every run); ``x0`` is one vector or ``(n_runs, n_features)``.  Keyword arguments
        and the fields of each returned OptimizeResult are the reference's
        (proximal_gradient.py:311-555); ``allvecs`` is not recorded."""
        from .proximal_gradient import _make_options, _message

        torch = _torch()
        if deprecated:
            warn("Using the deprecated option is not mathematically proven to converge. "
                 "Please consider using the recommended condition instead.", stacklevel=2)
        t0 = time.time()
        K, n = self.n_runs, self.n_features
        if nesterov_ratios is None:
            ab = np.tile(np.asarray(nesterov_ratio, dtype=np.float64), (K, 1))
        else:
            ab = np.ascontiguousarray(np.asarray(nesterov_ratios, dtype=np.float64).reshape(-1, 2))
            if len(ab) != K:
                raise ValueError(f"nesterov_ratios must have {K} rows")
        x0d, batched = self._x_batch(x0)
        cap = 0
        if return_all:
            cap = int(trace_capacity) if trace_capacity is not None else int(min(max_iter, 1 << 16))
        L = _lib.lib()
        while True:
            opts = _make_options(lr, tol, tol_internal, max_iter, 100000, max_backtrack_iter,
                                 False, decay_rate, nesterov, nesterov_ratio, deprecated,
                                 "reference", cap)
            xd = torch.empty(K, n, dtype=torch.float64, device=self.device)
            fun, lrs, err = np.empty(K), np.empty(K), np.empty(K)
            nit, status = np.zeros(K, dtype=np.int64), np.zeros(K, dtype=np.int32)
            allerrs = np.zeros((K, cap)) if cap else None
            allfuns = np.zeros((K, cap + 1)) if cap else None
            p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
            with torch.cuda.device(self.device):
                self._use_current_stream()
                if not self.distributed:
                    _lib.check(L.zf_lasso_multi_solve(
                        self._h, C.byref(opts), C.c_void_p(x0d.data_ptr()), int(batched), p(ab),
                        C.c_void_p(xd.data_ptr()), p(fun), p(nit), p(status), p(lrs), p(err),
                        p(allerrs), p(allfuns)))
                else:
                    if cap:
                        raise NotImplementedError("return_all is not available on the row-sharded path")
                    from .distributed import run_device_lasso, run_split_lasso

                    h = self._h

                    class _Ops:            # host-decided rounds (ZF_LASSO_HOSTLOOP=1)
                        def begin(self_):
                            _lib.check(L.zf_lasso_multi_begin(h, C.byref(opts),
                                                              C.c_void_p(x0d.data_ptr()),
                                                              int(batched), p(ab)))

                        def grad(self_, which):
                            _lib.check(L.zf_lasso_multi_grad(h, int(which)))

                        def step(self_):
                            nxt = C.c_int32(0)
                            _lib.check(L.zf_lasso_multi_step(h, C.byref(nxt)))
                            return nxt.value

                        def finish(self_):
                            _lib.check(L.zf_lasso_multi_finish(h, C.c_void_p(xd.data_ptr()), p(fun),
                                                               p(nit), p(status), p(lrs), p(err)))

                    class _DevOps:         # device-decided rounds: the host only enqueues and polls
                        def begin(self_):
                            _lib.check(L.zf_lasso_multi_dev_begin(h, C.byref(opts),
                                                                  C.c_void_p(x0d.data_ptr()),
                                                                  int(batched), p(ab)))

                        def stage(self_, k):
                            _lib.check(L.zf_lasso_multi_dev_stage(h, int(k)))

                        def needs_feval(self_):
                            return bool(L.zf_lasso_multi_dev_needs_feval(h))

                        def snapshot(self_, slot):
                            _lib.check(L.zf_lasso_multi_dev_poll(h, int(slot), 0, None))

                        def wait(self_, slot):
                            done = C.c_int32(0)
                            _lib.check(L.zf_lasso_multi_dev_poll(h, int(slot), 1, C.byref(done)))
                            return bool(done.value)

                        def finish(self_):
                            _lib.check(L.zf_lasso_multi_dev_finish(h, C.c_void_p(xd.data_ptr()),
                                                                   p(fun), p(nit), p(status), p(lrs),
                                                                   p(err)))

                    if os.environ.get("ZF_LASSO_HOSTLOOP", "0") == "1":
                        run_split_lasso(_Ops(), self._allreduce)
                    else:
                        run_device_lasso(_DevOps(), self._allreduce, self._allreduce_ss, chunk=8)
            if cap and int(nit.max()) > cap and trace_capacity is None:
                cap = int(nit.max())
                continue
            break
        elapsed = time.time() - t0
        xh = None if return_device else xd.cpu().numpy()
        out = []
        for k in range(K):
            st, it = int(status[k]), int(nit[k])
            res = OptimizeResult(
                x=xd[k] if return_device else xh[k], fun=float(fun[k]), nit=it, success=st == 1,
                status=st, message=_message(st), time=elapsed, lr=float(lrs[k]),
                err=float(err[k]), nesterov_ratio=(float(ab[k, 0]), float(ab[k, 1])),
                allvecs=None, allfuns=None, allerrs=None)
            if return_all:
                res.allerrs = list(allerrs[k, :it])
                res.allfuns = list(allfuns[k, :it + 1])
            out.append(res)
        return out
