"""ctypes binding of include/zfista_b200.h (the C-ABI drop-in boundary).

The library is the product: if ``libzfista_b200.so`` is missing, or no CUDA device is
visible when a compute entry point is called, the call fails loudly.  There is no
CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

ZF_MAX_OBJECTIVES = 4
_LIB_PATH = os.environ.get("ZFISTA_B200_LIB") or os.path.join(
    os.path.dirname(os.path.abspath(__file__)), "libzfista_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)
c_int32_p = C.POINTER(C.c_int32)


class ZfProblem(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("n_features", C.c_int32), ("n_objectives", C.c_int32),
        ("has_l1", C.c_int32),
        ("l1_ratios", C.c_double * ZF_MAX_OBJECTIVES),
        ("l1_shifts", C.c_double * ZF_MAX_OBJECTIVES),
        ("has_bounds", C.c_int32), ("bounds_are_arrays", C.c_int32),
        ("lower", C.c_double), ("upper", C.c_double),
        ("lower_v", C.c_void_p), ("upper_v", C.c_void_p),
        ("A", C.c_void_p), ("b", C.c_void_p),
        ("n_rows", C.c_int32), ("reserved", C.c_int32),
        ("scale", C.c_double), ("l1", C.c_double),
    ]


class ZfOptions(C.Structure):
    _fields_ = [
        ("lr", C.c_double), ("tol", C.c_double), ("tol_internal", C.c_double),
        ("max_iter", C.c_int64),
        ("max_iter_internal", C.c_int32), ("max_backtrack_iter", C.c_int32),
        ("warm_start", C.c_int32), ("nesterov", C.c_int32),
        ("decay_rate", C.c_double), ("nesterov_a", C.c_double), ("nesterov_b", C.c_double),
        ("deprecated", C.c_int32), ("dual_solver", C.c_int32),
        ("trace_capacity", C.c_int32), ("reserved", C.c_int32),
    ]


class ZfResult(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("fun", C.c_void_p), ("nit", C.c_void_p), ("status", C.c_void_p),
        ("lr", C.c_void_p), ("nfev", C.c_void_p), ("n_dual", C.c_void_p), ("err", C.c_void_p),
        ("allerrs", C.c_void_p), ("allfuns", C.c_void_p), ("allvecs", C.c_void_p),
        ("trace_offsets", C.c_void_p),
    ]


# every symbol include/zfista_b200.h declares (tests check the .so exports all of them)
EXPORTED_SYMBOLS = [
    "zf_abi_version", "zf_last_error", "zf_device_count", "zf_default_options",
    "zf_solve_batched_device", "zf_solve_batched_host", "zf_solve_subproblem_host",
    "zf_problem_eval_host",
    "zf_lasso_create", "zf_lasso_destroy", "zf_lasso_solve", "zf_lasso_begin",
    "zf_lasso_grad", "zf_lasso_partial", "zf_lasso_step", "zf_lasso_finish",
    "zf_lasso_gradient_device", "zf_lasso_passes",
    "zf_lasso_set_stream", "zf_lasso_dev_begin", "zf_lasso_dev_stage", "zf_lasso_dev_needs_feval",
    "zf_lasso_dev_slots", "zf_lasso_dev_poll", "zf_lasso_dev_finish",
    "zf_lasso_p2p_export", "zf_lasso_p2p_attach", "zf_lasso_p2p_active", "zf_lasso_set_allvecs",
    "zf_lasso_multi_create", "zf_lasso_multi_destroy", "zf_lasso_multi_set_stream",
    "zf_lasso_multi_solve",
    "zf_lasso_multi_begin", "zf_lasso_multi_grad", "zf_lasso_multi_partial",
    "zf_lasso_multi_step", "zf_lasso_multi_finish", "zf_lasso_multi_gradient_device",
    "zf_lasso_multi_pass_device",
    "zf_lasso_multi_dev_begin", "zf_lasso_multi_dev_stage", "zf_lasso_multi_dev_needs_feval",
    "zf_lasso_multi_dev_poll", "zf_lasso_multi_dev_finish", "zf_lasso_multi_layout",
    "zf_deblur_create", "zf_deblur_destroy", "zf_deblur_solve_host", "zf_deblur_solve_device",
    "zf_deblur_eval_host",
]


class ZfError(RuntimeError):
    """A zfista_b200 C-ABI call failed (code, message from zf_last_error())."""

    def __init__(self, code, message):
        super().__init__(f"zfista_b200 error {code}: {message}")
        self.code = code


_lib = None


def lib():
    """Load (once) and return the C-ABI library; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise ImportError(
            f"{_LIB_PATH} not found: build the CUDA library first "
            "(python -m zfista_b200.build, or __graft_entry__.build()). "
            "zfista_b200 has no CPU fallback.")
    L = C.CDLL(_LIB_PATH)
    L.zf_abi_version.restype = C.c_int
    L.zf_last_error.restype = C.c_char_p
    L.zf_device_count.restype = C.c_int
    L.zf_launch_count.restype = C.c_int64
    L.zf_default_options.argtypes = [C.POINTER(ZfOptions)]
    L.zf_default_options.restype = None
    L.zf_solve_batched_device.argtypes = [
        C.POINTER(ZfProblem), C.POINTER(ZfOptions), C.c_int64, C.c_void_p, C.c_void_p,
        C.POINTER(ZfResult), C.c_void_p]
    L.zf_solve_batched_host.argtypes = [
        C.POINTER(ZfProblem), C.POINTER(ZfOptions), C.c_int64, C.c_void_p, C.c_void_p,
        C.POINTER(ZfResult)]
    L.zf_solve_subproblem_host.argtypes = [
        C.POINTER(ZfProblem), C.POINTER(ZfOptions), C.c_int64, C.c_void_p, C.c_void_p,
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.zf_problem_eval_host.argtypes = [
        C.POINTER(ZfProblem), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
        C.c_void_p, C.c_void_p]
    for name in ("zf_solve_batched_device", "zf_solve_batched_host",
                 "zf_solve_subproblem_host", "zf_problem_eval_host"):
        getattr(L, name).restype = C.c_int
    # large-n LASSO handle API
    _bind_lasso(L)
    _lib = L
    return L


def _bind_lasso(L):
    L.zf_lasso_create.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_int64,
                                  C.c_int64, C.c_double, C.c_double, C.c_void_p]
    L.zf_lasso_create.restype = C.c_int
    L.zf_lasso_destroy.argtypes = [C.c_void_p]
    L.zf_lasso_destroy.restype = None
    L.zf_lasso_solve.argtypes = [C.c_void_p, C.POINTER(ZfOptions), C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.zf_lasso_solve.restype = C.c_int
    L.zf_lasso_begin.argtypes = [C.c_void_p, C.POINTER(ZfOptions), C.c_void_p]
    L.zf_lasso_begin.restype = C.c_int
    L.zf_lasso_grad.argtypes = [C.c_void_p, C.c_int]
    L.zf_lasso_grad.restype = C.c_int
    L.zf_lasso_partial.argtypes = [C.c_void_p, c_int64_p]
    L.zf_lasso_partial.restype = C.c_void_p
    L.zf_lasso_step.argtypes = [C.c_void_p, c_int32_p]
    L.zf_lasso_step.restype = C.c_int
    L.zf_lasso_finish.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.zf_lasso_finish.restype = C.c_int
    L.zf_lasso_gradient_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.zf_lasso_gradient_device.restype = C.c_int
    L.zf_lasso_passes.argtypes = [C.c_void_p]
    L.zf_lasso_passes.restype = C.c_int
    V, I32, I64, D = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    # device-decided loop
    L.zf_lasso_set_stream.argtypes = [V, V]
    L.zf_lasso_dev_begin.argtypes = [V, C.POINTER(ZfOptions), V, I32, I32]
    L.zf_lasso_dev_stage.argtypes = [V, I32]
    L.zf_lasso_dev_needs_feval.argtypes = [V]
    L.zf_lasso_dev_slots.argtypes = [V, I32]
    L.zf_lasso_dev_poll.argtypes = [V, I32, I32, c_int32_p, c_int64_p]
    L.zf_lasso_dev_finish.argtypes = [V, V, V, V, V, V, V, V]
    L.zf_lasso_set_allvecs.argtypes = [V, V, I64]
    L.zf_lasso_set_allvecs.restype = C.c_int
    L.zf_lasso_p2p_export.argtypes = [V, V]
    L.zf_lasso_p2p_attach.argtypes = [V, I32, I32, V]
    L.zf_lasso_p2p_active.argtypes = [V]
    for name in ("zf_lasso_p2p_export", "zf_lasso_p2p_attach", "zf_lasso_p2p_active"):
        getattr(L, name).restype = C.c_int
    for name in ("zf_lasso_set_stream", "zf_lasso_dev_begin", "zf_lasso_dev_stage",
                 "zf_lasso_dev_needs_feval", "zf_lasso_dev_slots", "zf_lasso_dev_poll",
                 "zf_lasso_dev_finish"):
        getattr(L, name).restype = C.c_int
    # many runs sharing one A (FP64 tensor-core passes)
    L.zf_lasso_multi_create.argtypes = [C.POINTER(V), V, I64, I64, V, I32, I32, D, D, V]
    L.zf_lasso_multi_create.restype = C.c_int
    L.zf_lasso_multi_destroy.argtypes = [V]
    L.zf_lasso_multi_destroy.restype = None
    L.zf_lasso_multi_set_stream.argtypes = [V, V]
    L.zf_lasso_multi_set_stream.restype = C.c_int
    L.zf_lasso_multi_solve.argtypes = [V, C.POINTER(ZfOptions), V, I32, V, V, V, V, V, V, V, V, V]
    L.zf_lasso_multi_solve.restype = C.c_int
    L.zf_lasso_multi_begin.argtypes = [V, C.POINTER(ZfOptions), V, I32, V]
    L.zf_lasso_multi_begin.restype = C.c_int
    L.zf_lasso_multi_grad.argtypes = [V, C.c_int]
    L.zf_lasso_multi_grad.restype = C.c_int
    L.zf_lasso_multi_partial.argtypes = [V, c_int64_p]
    L.zf_lasso_multi_partial.restype = V
    L.zf_lasso_multi_step.argtypes = [V, c_int32_p]
    L.zf_lasso_multi_step.restype = C.c_int
    L.zf_lasso_multi_finish.argtypes = [V, V, V, V, V, V, V]
    L.zf_lasso_multi_finish.restype = C.c_int
    L.zf_lasso_multi_gradient_device.argtypes = [V, V, V, V]
    L.zf_lasso_multi_gradient_device.restype = C.c_int
    L.zf_lasso_multi_pass_device.argtypes = [V, V, C.c_int]
    L.zf_lasso_multi_pass_device.restype = C.c_int
    L.zf_lasso_multi_dev_begin.argtypes = [V, C.POINTER(ZfOptions), V, I32, V]
    L.zf_lasso_multi_dev_begin.restype = C.c_int
    L.zf_lasso_multi_dev_stage.argtypes = [V, I32]
    L.zf_lasso_multi_dev_stage.restype = C.c_int
    L.zf_lasso_multi_dev_needs_feval.argtypes = [V]
    L.zf_lasso_multi_dev_needs_feval.restype = C.c_int
    L.zf_lasso_multi_dev_poll.argtypes = [V, I32, I32, c_int32_p]
    L.zf_lasso_multi_dev_poll.restype = C.c_int
    L.zf_lasso_multi_dev_finish.argtypes = [V, V, V, V, V, V, V]
    L.zf_lasso_multi_dev_finish.restype = C.c_int
    L.zf_lasso_multi_layout.argtypes = [V, c_int64_p, c_int64_p]
    L.zf_lasso_multi_layout.restype = C.c_int
    # deblurring handle API
    L.zf_deblur_create.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_void_p,
                                   C.c_int32, C.c_void_p, C.c_double, C.c_int32, C.c_void_p]
    L.zf_deblur_create.restype = C.c_int
    L.zf_deblur_destroy.argtypes = [C.c_void_p]
    L.zf_deblur_destroy.restype = None
    for name in ("zf_deblur_solve_host", "zf_deblur_solve_device"):
        fn = getattr(L, name)
        fn.argtypes = [C.c_void_p, C.POINTER(ZfOptions), C.c_int64, C.c_void_p, C.c_int32,
                       C.c_void_p, C.POINTER(ZfResult)]
        fn.restype = C.c_int
    L.zf_deblur_eval_host.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p]
    L.zf_deblur_eval_host.restype = C.c_int


def check(code: int) -> None:
    if code != 0:
        raise ZfError(code, lib().zf_last_error().decode("utf-8", "replace"))


def launch_count() -> int:
    """Kernels launched by the library in this process (bench.py's gpu_launches)."""
    return int(lib().zf_launch_count())


def default_options() -> ZfOptions:
    o = ZfOptions()
    lib().zf_default_options(C.byref(o))
    return o
