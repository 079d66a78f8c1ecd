"""Interleaved A/B of the chunk-ring kernel with one and two rows per exchange step
(ZF_LASSO_RING_RPS, read when the handle is created), same A, same clock state: both handles are
created first, the GPU is warmed into its sustained state, then the two are timed in alternation.

    python profiles/ab_ring_rps.py [ROWSxCOLS[:cluster] ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from zfista_b200.lasso import DenseLasso  # noqa: E402

args = sys.argv[1:] or ["100000x20000", "100000x20000:4", "100000x20000:5", "30000x36000:6", "65536x16384:4"]
peak = 6552.0
if os.path.exists("MEASURED_PEAKS.json"):
    peak = float(json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", peak))
os.environ["ZF_LASSO_TUNE"] = "0"
for a in args:
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
    probs = {}
    for rps in (1, 2):
        os.environ["ZF_LASSO_RING_RPS"] = str(rps)
        probs[rps] = DenseLasso(A, b, 1e-3, scale=1.0 / (2 * rows))
    for _ in range(300):                       # ~1 s: out of the burst clocks
        probs[1].gradient(x)
    torch.cuda.synchronize()
    tot = {1: 0.0, 2: 0.0}
    rounds, reps = 8, 20
    for _ in range(rounds):
        for rps in (1, 2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                probs[rps].gradient(x)
            e1.record()
            torch.cuda.synchronize()
            tot[rps] += e0.elapsed_time(e1) / reps
    ms = {k: v / rounds for k, v in tot.items()}
    print(json.dumps({"rows": rows, "cols": cols, "cluster": cl or "static policy",
                      "ms": {k: round(v, 4) for k, v in ms.items()},
                      "frac": {k: round(8.0 * rows * cols / v / 1e6 / peak, 3) for k, v in ms.items()}}),
          flush=True)
    del probs, A, b, x
