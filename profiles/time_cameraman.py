"""us per round of 15 runs (256x256, 9x9 blur, fixed step) for each stencil form of the
deblurring tile kernel:  python profiles/time_cameraman.py [iters]"""
import os
import sys
import time
import warnings

import numpy as np

sys.path.insert(0, ".")
from bench import AB_GRID, synthetic_observation  # noqa: E402
from zfista_b200.deblur import HaarDeblurL1, gaussian_kernel, lipschitz_constant  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
kernel = gaussian_kernel(9, 4.0)
kernel /= kernel.sum()
obs = synthetic_observation(256, 256, kernel, seed=1)
L = lipschitz_constant(kernel)
for form in ("general", "sym", "sep"):
    os.environ["ZF_DEBLUR_FORM"] = form
    prob = HaarDeblurL1(obs, kernel, 2e-5)
    x0 = prob.dwt_array(obs)
    kw = dict(lr=1 / L, decay_rate=1, nesterov=True, tol=0.0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        prob.minimize_proximal_gradient_batched(x0, np.array(AB_GRID), max_iter=20, **kw)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            res = prob.minimize_proximal_gradient_batched(x0, np.array(AB_GRID), max_iter=iters, **kw)
            best = min(best, time.perf_counter() - t0)
    print(f"{form:8s} {1e6 * best / iters:7.2f} us per round of 15 runs = "
          f"{15 * iters / best:9.0f} FISTA it/s   F = {float(res[0].fun[0]):.12e}")
