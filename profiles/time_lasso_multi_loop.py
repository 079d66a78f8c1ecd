"""16 runs sharing A (65536x16384): sustained time of one gradient of all runs against the
lockstep FISTA rounds of zf_lasso_multi_solve (device-decided; ZF_LASSO_HOSTLOOP=1: host-decided).

    python profiles/time_lasso_multi_loop.py [rows cols runs iters]"""
import sys
import time
import warnings

import torch

sys.path.insert(0, ".")
from bench import AB_GRID  # noqa: E402
from zfista_b200.lasso import DenseLassoMulti  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
K = int(sys.argv[3]) if len(sys.argv) > 3 else 16
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 60
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(7)
A = torch.empty(rows, cols, dtype=torch.float64, device=dev)
chunk = max(1, (64 << 20) // (cols * 8))
for r0 in range(0, rows, chunk):
    A[r0:r0 + chunk] = torch.randn(min(chunk, rows - r0), cols, dtype=torch.float64, device=dev, generator=g)
w = torch.zeros(cols, dtype=torch.float64, device=dev)
w[:64] = 1.0
b = A @ w
prob = DenseLassoMulti(A, b, 1e-3, K, scale=1.0 / (2 * rows))
grid = [AB_GRID[k % len(AB_GRID)] for k in range(K)]
x = torch.zeros(cols, dtype=torch.float64, device=dev)
kw = dict(lr=0.5, decay_rate=1, nesterov=True, tol=0.0, return_device=True)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    X = torch.stack([r.x for r in prob.minimize_proximal_gradient_batched(x, grid, max_iter=5, **kw)])
    for _ in range(3):
        prob.gradient(X)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(40):
        prob.gradient(X)
    e1.record()
    e1.synchronize()
    gms = e0.elapsed_time(e1) / 40
    print(f"gradient of all {K} runs x40 (sustained): {gms:.4f} ms")
    ts = []
    for n in (10, 10, 10 + iters):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = prob.minimize_proximal_gradient_batched(x, grid, max_iter=n, **kw)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    per_round = 1e3 * (ts[2] - ts[1]) / iters
    print(f"lockstep round: {per_round:.4f} ms = {K / per_round * 1e3:.0f} run-iterations/s = "
          f"{gms / per_round:.3f} of the gradient-bound rate")
    t0 = time.perf_counter()
    res = prob.minimize_proximal_gradient_batched(x, grid, max_iter=20, nesterov=True, tol=0.0,
                                                  return_device=True)
    torch.cuda.synchronize()
    print(f"with the line search: {1e3 * (time.perf_counter() - t0) / 20:.4f} ms per round")
