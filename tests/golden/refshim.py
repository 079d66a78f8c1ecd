"""Import shim used ONLY to generate golden vectors from the real reference.

`/root/reference/zfista/problems.py` imports `jax` and `jaxopt`
(pyproject.toml dependency "jaxopt", unpinned; public releases 0.8.x), which are not
installed in this image and cannot be installed (no network).  The reference uses
exactly two functions from it (problems.py:8-9):

  * ``jaxopt.prox.prox_lasso(x, l1reg, scaling=1.0)``
        published definition: sign(x) * max(|x| - l1reg * scaling, 0)
  * ``jaxopt.projection.projection_box(x, (lower, upper))``
        published definition: clip(x, lower, upper)

and ``jax.config.update("jax_enable_x64", True)`` (fp64 everywhere).  This module
registers numpy restatements of those two published definitions under the same
module names so that the UNMODIFIED reference package can be imported from
/root/reference and run to produce fixtures.  It is never imported by the product
path, the tests, or bench.py -- only by tests/golden/make_golden.py.
"""
from __future__ import annotations

import sys
import types

import numpy as np


def install(reference_root: str = "/root/reference") -> None:
    if "jaxopt" in sys.modules:
        return
    jax = types.ModuleType("jax")
    jax.config = types.SimpleNamespace(update=lambda *a, **k: None)
    jaxopt = types.ModuleType("jaxopt")
    prox = types.ModuleType("jaxopt.prox")
    projection = types.ModuleType("jaxopt.projection")

    def prox_lasso(x, l1reg=None, scaling=1.0):
        if l1reg is None:
            l1reg = 1.0
        x = np.asarray(x, dtype=np.float64)
        return np.sign(x) * np.maximum(np.abs(x) - l1reg * scaling, 0.0)

    def projection_box(x, hyperparams):
        lower, upper = hyperparams
        return np.clip(np.asarray(x, dtype=np.float64), lower, upper)

    prox.prox_lasso = prox_lasso
    projection.projection_box = projection_box
    jaxopt.prox = prox
    jaxopt.projection = projection
    sys.modules["jax"] = jax
    sys.modules["jaxopt"] = jaxopt
    sys.modules["jaxopt.prox"] = prox
    sys.modules["jaxopt.projection"] = projection
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
