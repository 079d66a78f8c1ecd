"""Time one single-run LASSO gradient A^T(A x - b) (csrc/zf_lasso.cu) with CUDA events, L2 flushed
between launches; the kernel form follows the handle's policy or the ZF_LASSO_* overrides.

    [ZF_LASSO_RING=2] python profiles/time_lasso.py ROWS COLS [REPS]
"""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from zfista_b200.lasso import DenseLasso  # noqa: E402

rows, cols = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
peak = 6552.0
if os.path.exists("MEASURED_PEAKS.json"):
    peak = float(json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", peak))
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn(rows, cols, dtype=torch.float64, device="cuda", generator=g)
b = torch.randn(rows, dtype=torch.float64, device="cuda", generator=g)
prob = DenseLasso(A, b, 1e-3, scale=1.0 / (2 * rows))
x = torch.randn(cols, dtype=torch.float64, device="cuda", generator=g)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
grad, f = prob.gradient(x)
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    prob.gradient(x)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
med = sorted(ts)[len(ts) // 2]
env = {k: v for k, v in os.environ.items() if k.startswith("ZF_LASSO")}
print(json.dumps({"rows": rows, "cols": cols, "env": env, "passes": prob.hbm_passes_per_gradient(),
                  "ms_median": round(med, 4), "ms_best": round(min(ts), 4),
                  "one_pass_hbm_frac": round(8.0 * rows * cols / med / 1e6 / peak, 3),
                  "f": float(f), "gnorm": float(grad.norm())}))
