"""Small driver for profiling the dense-LASSO gradient kernels: rows x cols fp64, a few gradients."""
import sys

import torch

sys.path.insert(0, ".")
from zfista_b200.lasso import DenseLasso  # noqa: E402

rows, cols, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 3
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn(rows, cols, dtype=torch.float64, device="cuda", generator=g)
b = torch.randn(rows, dtype=torch.float64, device="cuda", generator=g)
prob = DenseLasso(A, b, 1e-3, scale=1.0 / (2 * rows))
x = torch.randn(cols, dtype=torch.float64, device="cuda", generator=g)
for _ in range(reps):
    grad, f = prob.gradient(x)
torch.cuda.synchronize()
print("f", float(f), "|grad|", float(grad.norm()))
