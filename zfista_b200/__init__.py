"""zfista_b200: B200-native proximal-gradient (ISTA / FISTA) hot path of zalgo3/zfista.

Public API mirrors the reference package::

    from zfista_b200 import minimize_proximal_gradient
    from zfista_b200.problems import JOS1, FDS, ...

Everything computes through the CUDA library ``libzfista_b200.so`` (C ABI in
``include/zfista_b200.h``).  Importing the package does not load it; the first call
does, and raises if it is missing or no GPU is visible -- there is no CPU fallback.
"""
from .proximal_gradient import (BatchResult, minimize_proximal_gradient,
                                minimize_proximal_gradient_batched, solve_subproblems)

__all__ = [
    "minimize_proximal_gradient",
    "minimize_proximal_gradient_batched",
    "solve_subproblems",
    "BatchResult",
]
__version__ = "0.1.0"
