"""Two batched solves of the headline workload (FDS n=100 +L1, FISTA, 1024 starts) for ncu."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zfista_b200.problems as zp  # noqa: E402

n = 100
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
prob = zp.FDS(n_features=n, l1_ratios=(np.arange(3) + 1) / n, l1_shifts=np.arange(3.0))
for seed in (1000, 1001):
    X0 = np.random.RandomState(seed).uniform(-2, 2, size=(S, n))
    br = prob.minimize_proximal_gradient_batched(X0, nesterov=True, tol_internal=1e-11,
                                                 max_iter=100000000)
    print(seed, int(br.nit.sum()), int(br.nit.max()), int((br.status == 1).sum()), br.time)
