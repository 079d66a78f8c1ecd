"""Driver for timing / profiling the shared-A multi-run LASSO passes (csrc/zf_lasso_multi.cu):
rows x cols fp64, n_runs runs; times pass 1 (R = A X - B), pass 2 (A^T R) and the whole gradient
with CUDA events and prints them against the measured HBM peak.

    python profiles/run_lasso_multi.py ROWS COLS RUNS [REPS] [--quiet]
"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from zfista_b200 import _lib  # noqa: E402
from zfista_b200.lasso import DenseLassoMulti  # noqa: E402

rows, cols, runs = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 and not sys.argv[4].startswith("-") else 5
peak = 6552.0
if os.path.exists("MEASURED_PEAKS.json"):
    peak = float(json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", peak))
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn(rows, cols, dtype=torch.float64, device="cuda", generator=g)
b = torch.randn(rows, dtype=torch.float64, device="cuda", generator=g)
prob = DenseLassoMulti(A, b, 1e-3, runs, scale=1.0 / (2 * rows))
X = torch.randn(runs, cols, dtype=torch.float64, device="cuda", generator=g)
L = _lib.lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts) // 2]


def p(which):
    _lib.check(L.zf_lasso_multi_pass_device(prob._h, C.c_void_p(X.data_ptr()), which))


grad, f = prob.gradient(X)          # warm-up + sanity
torch.cuda.synchronize()
bytes_a = 8.0 * rows * cols
out = {"rows": rows, "cols": cols, "runs": runs, "GiB": bytes_a / 2**30}
for name, fn in (("pass1_residual", lambda: p(0)), ("pass2_atr", lambda: p(1)),
                 ("gradient", lambda: prob.gradient(X))):
    best, med = timed(fn)
    n_pass = 2 if name == "gradient" else 1
    out[name] = {"ms_best": round(best, 4), "ms_median": round(med, 4),
                 "GBps": round(n_pass * bytes_a / med / 1e6, 1),
                 "hbm_frac": round(n_pass * bytes_a / med / 1e6 / peak, 3),
                 "fp64_tflops": round(n_pass * 2.0 * rows * cols * prob.n_runs / med / 1e9, 2)}
out["per_run_gradient_ms"] = round(out["gradient"]["ms_median"] / runs, 4)
print(json.dumps(out))
print("f0", float(f[0]), "|grad0|", float(grad[0].norm()), file=sys.stderr)
