"""Sustained gradient-pass time of the LASSO kernel the library picks for each shape (create-time
probe on unless ZF_LASSO_TUNE=0): 300 warm-up passes (~1 s: out of the burst clocks), then
8 x 20 passes timed with CUDA events.

    python profiles/time_ring_sustained.py ROWSxCOLS[:cluster] ..."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from zfista_b200.lasso import DenseLasso  # noqa: E402

peak = 6552.0
if os.path.exists("MEASURED_PEAKS.json"):
    peak = float(json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", peak))
for a in sys.argv[1:]:
    shape, _, cl = a.partition(":")
    rows, cols = map(int, shape.split("x"))
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.empty(rows, cols, dtype=torch.float64, device="cuda")
    chunk = max(1, (64 << 20) // (cols * 8))
    for r0 in range(0, rows, chunk):
        A[r0:r0 + chunk] = torch.randn(min(chunk, rows - r0), cols, dtype=torch.float64, device="cuda",
                                       generator=g)
    b = torch.randn(rows, dtype=torch.float64, device="cuda", generator=g)
    x = torch.randn(cols, dtype=torch.float64, device="cuda", generator=g)
    if cl:
        os.environ["ZF_LASSO_RING"] = cl
    else:
        os.environ.pop("ZF_LASSO_RING", None)
    prob = DenseLasso(A, b, 1e-3, scale=1.0 / (2 * rows))
    grad, _ = prob.gradient(x)
    err = None
    if rows * cols <= 2_500_000_000:
        ref = (A.T @ (A @ x - b)) / rows
        err = float((grad - ref).abs().max() / ref.abs().max())
    for _ in range(300):
        prob.gradient(x)
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            prob.gradient(x)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1) / 20
    ms = tot / 8
    print(json.dumps({"rows": rows, "cols": cols, "cluster": cl or "library's choice",
                      "ms": round(ms, 4), "frac": round(8.0 * rows * cols / ms / 1e6 / peak, 3),
                      "rel_err_vs_torch": err}), flush=True)
    del prob, A, b, x
