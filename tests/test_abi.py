"""CPU tests of the drop-in boundary: the C-ABI library builds, loads, and exports every
symbol include/zfista_b200.h declares; the ctypes structs match the header's layout; and
the product path refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "zfista_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built_lib):
    from zfista_b200 import _lib

    declared = _declared_functions()
    assert len(declared) >= 17
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True,
                         text=True, check=True).stdout
    for name in declared:
        assert re.search(rf"\bT {name}\b", out), f"{name} is not an exported text symbol"
    assert L.zf_abi_version() == 1


def test_struct_layout_matches_header(built_lib, tmp_path):
    """sizeof / offsetof as the C compiler sees the header == the ctypes mirror."""
    from zfista_b200 import _lib

    probe = tmp_path / "probe.c"
    probe.write_text('''
#include <stddef.h>
#include <stdio.h>
#include "zfista_b200.h"
int main(void) {
  printf("%zu %zu %zu\\n", sizeof(zf_problem), sizeof(zf_options), sizeof(zf_result));
  printf("%zu %zu %zu %zu %zu\\n", offsetof(zf_problem, l1_shifts), offsetof(zf_problem, lower_v),
         offsetof(zf_problem, A), offsetof(zf_problem, scale), offsetof(zf_problem, l1));
  printf("%zu %zu %zu %zu\\n", offsetof(zf_options, max_iter), offsetof(zf_options, decay_rate),
         offsetof(zf_options, deprecated), offsetof(zf_options, trace_capacity));
  printf("%zu %zu %zu\\n", offsetof(zf_result, lr), offsetof(zf_result, allvecs),
         offsetof(zf_result, trace_offsets));
  return 0;
}
''')
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(probe), "-o", str(exe)],
                   check=True)
    vals = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True,
                                           check=True).stdout.split()]
    P, O, R = _lib.ZfProblem, _lib.ZfOptions, _lib.ZfResult
    assert vals == [
        C.sizeof(P), C.sizeof(O), C.sizeof(R),
        P.l1_shifts.offset, P.lower_v.offset, P.A.offset, P.scale.offset, P.l1.offset,
        O.max_iter.offset, O.decay_rate.offset, O.deprecated.offset, O.trace_capacity.offset,
        R.lr.offset, R.allvecs.offset, R.trace_offsets.offset,
    ]


def test_default_options_are_the_reference_defaults(built_lib):
    """proximal_gradient.py:311-331."""
    from zfista_b200 import _lib

    o = _lib.default_options()
    assert (o.lr, o.tol, o.tol_internal) == (1.0, 1e-5, 1e-12)
    assert (o.max_iter, o.max_iter_internal, o.max_backtrack_iter) == (1000000, 100000, 100)
    assert (o.warm_start, o.nesterov, o.deprecated) == (0, 0, 0)
    assert (o.decay_rate, o.nesterov_a, o.nesterov_b) == (0.5, 0.0, 0.25)


def test_problem_descriptors_and_names(built_lib):
    """Constructor defaults, names and validation mirror zfista/problems.py."""
    import zfista_b200.problems as zp

    assert zp.JOS1().name == "JOS1_n_5"
    p = zp.JOS1(n_features=5, l1_ratios=[0.2, 0.1], l1_shifts=[0, 1])
    assert p.name == "JOS1_n_5_l1_ratios_0.2_0.1_l1_shifts_0_1"
    assert zp.SD().bounds == (1e-6, np.inf) and zp.SD().n_features == 4
    assert zp.ZDT1().n_features == 30 and zp.FDS().n_objectives == 3
    assert zp.TRIDIA().n_features == 3 and zp.LinearFunctionRank1().n_objectives == 4
    d, keep = zp.FDS(n_features=7, bounds=(np.zeros(7), np.inf)).descriptor()
    assert d.kind == 2 and d.n_features == 7 and d.has_bounds == 1 and d.bounds_are_arrays == 1
    with pytest.raises(ValueError):
        zp.JOS1(l1_ratios=[0.1]).descriptor()
    with pytest.raises(NotImplementedError):
        zp.LinearFunctionRank1(n_objectives=6).descriptor()


def test_callables_that_are_not_device_functors_are_rejected(built_lib):
    from zfista_b200 import minimize_proximal_gradient

    with pytest.raises(TypeError, match="no CPU fallback"):
        minimize_proximal_gradient(lambda x: x @ x, lambda x: 0.0, lambda x: 2 * x,
                                   lambda w, x: x, np.zeros(3))


def test_no_cpu_fallback_without_a_device(built_lib):
    """On a box without a GPU every compute entry point fails loudly."""
    from zfista_b200 import _lib
    import zfista_b200.problems as zp

    if _lib.lib().zf_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(_lib.ZfError, match="no CPU fallback"):
        zp.JOS1().minimize_proximal_gradient(np.zeros(5))
    with pytest.raises(_lib.ZfError):
        zp.JOS1().f(np.zeros(5))
    from zfista_b200.lasso import DenseLasso

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        DenseLasso(np.eye(3), np.ones(3), 0.1)


def test_shared_a_entry_points_validate_and_refuse_without_a_device(built_lib):
    """zf_lasso_multi_*: argument validation happens before any device work, and on a box
    without a GPU creation fails with ZF_ERR_CUDA (no compute call is made here)."""
    from zfista_b200 import _lib
    from zfista_b200.lasso import DenseLassoMulti

    L = _lib.lib()
    h = C.c_void_p()
    fake = C.c_void_p(0x1000)          # never dereferenced: validation comes first

    def create(rows, cols, runs, ptr=fake):
        return L.zf_lasso_multi_create(C.byref(h), ptr, rows, cols, fake, 0, runs, 1.0, 0.1, None)

    assert create(4, 4, 2, None) == -1 and b"NULL" in L.zf_last_error()
    assert create(0, 4, 2) == -1
    assert create(4, 4, 0) == -1 and create(4, 4, 33) == -1
    assert create(4, 3, 2) == -3 and b"even" in L.zf_last_error()            # odd n_cols
    assert create(4, 4, 2, C.c_void_p(0x1008)) == -3                          # misaligned A
    assert L.zf_lasso_multi_step(None, None) == -1
    assert L.zf_lasso_multi_grad(None, 0) == -1
    assert L.zf_lasso_multi_finish(None, None, None, None, None, None, None) == -1
    assert L.zf_lasso_multi_partial(None, None) is None
    L.zf_lasso_multi_destroy(None)                                             # no-op
    if L.zf_device_count() > 0:
        pytest.skip("a GPU is visible: the no-device half does not apply")
    assert create(4, 4, 2) == -2 and b"no CPU fallback" in L.zf_last_error()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        DenseLassoMulti(np.eye(4), np.ones(4), 0.1, 2)
