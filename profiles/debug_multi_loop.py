import sys, warnings, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import helpers
from zfista_b200.lasso import DenseLassoMulti


def _dataset(seed, n_rows, n_cols, n_runs, batched_b):
    rng = np.random.RandomState(seed)
    A = rng.standard_normal((n_rows, n_cols))
    w = np.zeros((n_runs if batched_b else 1, n_cols))
    w[:, :10] = rng.standard_normal((len(w), 10))
    b = w @ A.T + 0.01 * rng.standard_normal((len(w), n_rows))
    X0 = rng.standard_normal((n_runs, n_cols)) * 0.1
    return A, (b if batched_b else b[0]), X0

n_rows, n_cols = 310, 144
grid = helpers.AB_GRID[:11]; K = len(grid)
A, b, X0 = _dataset(77, n_rows, n_cols, K, True)
scale, l1 = 1 / (2 * n_rows), 0.05
prob = DenseLassoMulti(A, b, l1, K, scale=scale)
lip = 2 * scale * np.linalg.norm(A, 2) ** 2
cases = [dict(nesterov=True), dict(nesterov=False, max_iter=30),
         dict(nesterov=True, lr=1 / lip, decay_rate=1, max_iter=250),
         dict(nesterov=True, lr=1 / lip, decay_rate=1, max_iter=19, tol=0.0),
         dict(nesterov=True, deprecated=True, lr=8.0)]
out = {}
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    for ci, opts in enumerate(cases):
        got = prob.minimize_proximal_gradient_batched(X0, grid, **opts)
        tr = prob.minimize_proximal_gradient_batched(X0, grid, return_all=True, **opts)
        for k in range(K):
            out[f"c{ci}_x{k}"] = got[k].x; out[f"c{ci}_e{k}"] = np.array(tr[k].allerrs); out[f"c{ci}_f{k}"] = np.array(tr[k].allfuns)
            out[f"c{ci}_xt{k}"] = tr[k].x; out[f"c{ci}_lr{k}"] = got[k].lr
np.savez(sys.argv[1], **out)
