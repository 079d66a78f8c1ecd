"""Small driver for profiling the deblurring kernels: 15 (a, b) runs, 256x256, N iterations."""
import sys
import warnings

import numpy as np

sys.path.insert(0, ".")
from bench import synthetic_observation  # noqa: E402
from zfista_b200.deblur import HaarDeblurL1, gaussian_kernel, lipschitz_constant  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 24
kernel = gaussian_kernel(9, 4.0)
kernel /= kernel.sum()
obs = synthetic_observation(256, 256, kernel, seed=1)
prob = HaarDeblurL1(obs, kernel, 2e-5)
x0 = prob.dwt_array(obs)
L = lipschitz_constant(kernel)
grid = [(0.0, 0.0), (0.0, 1 / 8), (0.0, 1 / 4), (1 / 6, 1 / 144), (1 / 6, 37 / 288), (1 / 6, 1 / 4),
        (1 / 4, 1 / 64), (1 / 4, 17 / 128), (1 / 4, 1 / 4), (1 / 2, 1 / 16), (1 / 2, 5 / 32),
        (1 / 2, 1 / 4), (3 / 4, 9 / 64), (3 / 4, 25 / 128), (3 / 4, 1 / 4)]
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    res = prob.minimize_proximal_gradient_batched(x0, np.array(grid), lr=1 / L, decay_rate=1,
                                                  nesterov=True, max_iter=iters, tol=0.0)
print("nit", [r.nit for r in res], "F", float(res[0].fun[0]))
