"""Small end-to-end runs of every kernel family for compute-sanitizer (one tool per call):
    compute-sanitizer --tool memcheck python profiles/sanitize_small.py
"""
import os
import sys
import warnings

import numpy as np

sys.path.insert(0, ".")
warnings.simplefilter("ignore")
which = sys.argv[1] if len(sys.argv) > 1 else "all"

if which in ("all", "batched"):
    import zfista_b200.problems as zp

    rng = np.random.RandomState(0)
    for prob, lo, hi in [(zp.JOS1(n_features=50, l1_ratios=(0.02, 0.01), l1_shifts=(0, 1)), -2, 4),
                         (zp.FDS(n_features=33, l1_ratios=(0.1, 0.2, 0.3), l1_shifts=(0, 1, 2)), -2, 2),
                         (zp.SD(), 1.5, 3), (zp.LinearFunctionRank1(n_features=30), -1, 1)]:
        X0 = rng.uniform(lo, hi, size=(37, prob.n_features))
        br = prob.minimize_proximal_gradient_batched(X0, nesterov=True, tol_internal=1e-11,
                                                     max_iter=60, return_all=True)
        print(type(prob).__name__, "nit", br.nit[:4], flush=True)

if which in ("all", "lasso"):
    import torch
    from zfista_b200.lasso import DenseLasso

    for env, shape in [({"ZF_LASSO_FUSED": "0"}, (700, 1030)), ({}, (700, 2050)),
                       ({"ZF_LASSO_RING": "2"}, (701, 3000)), ({"ZF_LASSO_RING": "3"}, (900, 20000)),
                       ({"ZF_LASSO_RING": "1"}, (1000, 5000)), ({}, (50, 201))]:
        for k in ("ZF_LASSO_FUSED", "ZF_LASSO_RING"):
            os.environ.pop(k, None)
        os.environ.update(env)
        rows, cols = shape
        g = torch.Generator(device="cuda").manual_seed(1)
        A = torch.randn(rows, cols, dtype=torch.float64, device="cuda", generator=g)
        b = torch.randn(rows, dtype=torch.float64, device="cuda", generator=g)
        prob = DenseLasso(A, b, 0.01, scale=1 / (2 * rows))
        x = torch.randn(cols, dtype=torch.float64, device="cuda", generator=g)
        grad, f = prob.gradient(x)
        ref = (A.T @ (A @ x - b)) / rows
        err = float((grad - ref).abs().max() / ref.abs().max())
        res = prob.minimize_proximal_gradient(np.zeros(cols), nesterov=True, max_iter=8)
        print(env, shape, "passes", prob.hbm_passes_per_gradient(), "grad rel err %.1e" % err,
              "nit", res.nit, flush=True)
        assert err < 1e-12

if which in ("all", "deblur"):
    from bench import synthetic_observation
    from zfista_b200.deblur import HaarDeblurL1, gaussian_kernel, lipschitz_constant

    for shape, ks in [((70, 34), 7), ((64, 96), 9), ((20, 36), 3)]:
        kernel = gaussian_kernel(ks, ks / 3.0)
        kernel /= kernel.sum()
        obs = synthetic_observation(shape[0], shape[1], kernel, seed=2)
        prob = HaarDeblurL1(obs, kernel, 1e-4, max_runs=4)
        x0 = prob.dwt_array(obs)
        L = lipschitz_constant(kernel)
        r1 = prob.minimize_proximal_gradient_batched(x0, [(0, 0.25), (0.25, 0.25)], lr=1 / L,
                                                     decay_rate=1, max_iter=20)
        r2 = prob.minimize_proximal_gradient_batched(x0, [(0, 0.25)], max_iter=10, return_all=True)
        print(shape, ks, "nit", r1[0].nit, r2[0].nit, "F", float(r1[0].fun[0]), flush=True)
print("SANITIZE_RUN_OK")
