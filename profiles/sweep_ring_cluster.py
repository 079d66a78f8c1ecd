"""Sustained gradient-pass time of the chunk-ring kernel for every cluster size that can hold a
row slice (ZF_LASSO_RING=c), per shape; checks each result against torch.

    python profiles/sweep_ring_cluster.py [ROWSxCOLS ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from zfista_b200.lasso import DenseLasso  # noqa: E402

shapes = [tuple(map(int, a.split("x"))) for a in sys.argv[1:]] or [
    (200000, 20000), (100000, 20000), (25000, 20000), (80000, 24000), (40000, 30000),
    (30000, 36000), (20000, 40000), (30000, 50000), (20000, 60000)]
peak = 6552.0
if os.path.exists("MEASURED_PEAKS.json"):
    peak = float(json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", peak))
reps = 30
for rows, cols in shapes:
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.empty(rows, cols, dtype=torch.float64, device="cuda")
    chunk = max(1, (64 << 20) // (cols * 8))
    for r0 in range(0, rows, chunk):
        A[r0:r0 + chunk] = torch.randn(min(chunk, rows - r0), cols, dtype=torch.float64,
                                       device="cuda", generator=g)
    b = torch.randn(rows, dtype=torch.float64, device="cuda", generator=g)
    x = torch.randn(cols, dtype=torch.float64, device="cuda", generator=g)
    ref = None
    if rows * cols <= 2_500_000_000:
        ref = (A.T @ (A @ x - b)) / rows
    for c in ("policy", 1, 2, 3, 4, 5, 6, 7, 8):
        if c == "policy":
            os.environ.pop("ZF_LASSO_RING", None)
        else:
            if (cols // 2 + c - 1) // c > 5 * 1024 or (cols // 2 + c - 1) // c <= 2 * 1024 and c > 1 \
                    and (cols // 2 + c - 2) // (c - 1) <= 2 * 1024:
                continue
            os.environ["ZF_LASSO_RING"] = str(c)
        prob = DenseLasso(A, b, 1e-3, scale=1.0 / (2 * rows))
        if prob.hbm_passes_per_gradient() != 1:
            continue
        grad, f = prob.gradient(x)
        err = None
        if ref is not None:
            err = float((grad - ref).abs().max() / ref.abs().max())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            prob.gradient(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(json.dumps({"rows": rows, "cols": cols, "cluster": c, "ms": round(ms, 4),
                          "frac": round(8.0 * rows * cols / ms / 1e6 / peak, 3),
                          "rel_err_vs_torch": err}), flush=True)
        del prob
    del A, b, x, ref
    torch.cuda.empty_cache()
