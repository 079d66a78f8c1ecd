"""Small driver for profiling the device-decided LASSO loop: a short fixed-step FISTA solve and a
short line-search solve (rows x cols fp64)."""
import sys
import warnings

import torch

sys.path.insert(0, ".")
from zfista_b200.lasso import DenseLasso  # noqa: E402

rows, cols = int(sys.argv[1]), int(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 12
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn(rows, cols, dtype=torch.float64, device="cuda", generator=g)
w = torch.zeros(cols, dtype=torch.float64, device="cuda")
w[:32] = 1.0
b = A @ w
prob = DenseLasso(A, b, 1e-3, scale=1.0 / (2 * rows))
x = torch.zeros(cols, dtype=torch.float64, device="cuda")
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    r1 = prob.minimize_proximal_gradient(x, lr=0.5, decay_rate=1, nesterov=True, tol=0.0, max_iter=iters)
    r2 = prob.minimize_proximal_gradient(x, nesterov=True, tol=0.0, max_iter=iters)
print("fixed", r1.nit, float(r1.fun), "line search", r2.nit, float(r2.fun))
