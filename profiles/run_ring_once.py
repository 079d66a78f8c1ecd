"""A few gradient passes of the chunk-ring kernel at one shape (for ncu captures).

    python profiles/run_ring_once.py ROWSxCOLS [passes]"""
import sys

import torch

sys.path.insert(0, ".")
from zfista_b200.lasso import DenseLasso  # noqa: E402

rows, cols = map(int, sys.argv[1].split("x"))
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 4
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.empty(rows, cols, dtype=torch.float64, device="cuda")
chunk = max(1, (64 << 20) // (cols * 8))
for r0 in range(0, rows, chunk):
    A[r0:r0 + chunk] = torch.randn(min(chunk, rows - r0), cols, dtype=torch.float64, device="cuda", generator=g)
b = torch.randn(rows, dtype=torch.float64, device="cuda", generator=g)
x = torch.randn(cols, dtype=torch.float64, device="cuda", generator=g)
prob = DenseLasso(A, b, 1e-3, scale=1.0 / (2 * rows))
for _ in range(passes):
    prob.gradient(x)
torch.cuda.synchronize()
print("ok")
