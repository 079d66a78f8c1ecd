"""Per-iteration cost of the fixed-step LASSO loop against the gradient pass alone, both
SUSTAINED (hundreds of back-to-back launches, so the clocks are the same in both):

    python profiles/time_lasso_loop.py [rows cols iters]

Prints ms per gradient (zf_lasso_gradient_device back to back), ms per FISTA iteration of
zf_lasso_solve (device-decided loop; ZF_LASSO_HOSTLOOP=1: host-decided), and their ratio."""
import os
import subprocess
import sys
import threading
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zfista_b200.lasso import DenseLasso  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 300
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(7)
A = torch.empty(rows, cols, dtype=torch.float64, device=dev)
chunk = max(1, (64 << 20) // (cols * 8))
for r0 in range(0, rows, chunk):
    r1 = min(rows, r0 + chunk)
    A[r0:r1] = torch.randn(r1 - r0, cols, dtype=torch.float64, device=dev, generator=g)
w = torch.zeros(cols, dtype=torch.float64, device=dev)
w[:64] = 1.0
b = A @ w
prob = DenseLasso(A, b, l1_ratio=1e-3, scale=1.0 / (2 * rows))
x = torch.zeros(cols, dtype=torch.float64, device=dev)
stream = torch.cuda.current_stream()

clk = []
stop = False


def sample():
    p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw",
                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                         text=True)
    while not stop:
        ln = p.stdout.readline()
        if ln:
            clk.append((time.time(), ln.strip()))
    p.terminate()


th = threading.Thread(target=sample, daemon=True)
th.start()


def clocks_between(t0, t1):
    v = [float(s.split(",")[0]) for t, s in clk if t0 <= t <= t1]
    return (np.median(v), len(v)) if v else (None, 0)


for reps in (5, iters):
    for _ in range(3):
        prob.gradient(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record(stream)
    for _ in range(reps):
        prob.gradient(x)
    e1.record(stream)
    e1.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"gradient x{reps}: {ms:.4f} ms each, {rows * cols * 8 / ms / 1e6:.0f} GB/s, "
          f"clocks {clocks_between(t0, time.time())}")
    grad_ms = ms
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    for n in (iters, iters, 2 * iters):
        kw = dict(lr=0.5, decay_rate=1, nesterov=True, tol=0.0, return_device=True)
        prob.minimize_proximal_gradient(x, max_iter=3, **kw)
        torch.cuda.synchronize()
        t0 = time.time()
        res = prob.minimize_proximal_gradient(x, max_iter=n, **kw)
        torch.cuda.synchronize()
        dt = time.time() - t0
        print(f"fista x{n}: {1e3 * dt / res.nit:.4f} ms per iteration = {res.nit / dt:.1f} it/s; "
              f"x gradient-bound rate {grad_ms / (1e3 * dt / res.nit):.3f}; clocks {clocks_between(t0, time.time())}")
    # with the line search (two passes over A per trial)
    t0 = time.time()
    res = prob.minimize_proximal_gradient(x, max_iter=iters // 2, nesterov=True, tol=0.0, return_device=True)
    torch.cuda.synchronize()
    dt = time.time() - t0
    print(f"fista+line search x{res.nit}: {1e3 * dt / res.nit:.4f} ms per iteration")
stop = True

# ---- per-stage breakdown of the device-decided loop (CUDA events between the stages)
import ctypes as C  # noqa: E402

from zfista_b200 import _lib  # noqa: E402
from zfista_b200.proximal_gradient import _make_options  # noqa: E402

L = _lib.lib()
opts = _make_options(0.5, 0.0, 1e-12, 10 ** 6, 100000, 100, False, 1.0, True, (0, 0.25), False,
                     "reference", 0)
prob._use_current_stream()
_lib.check(L.zf_lasso_dev_begin(prob._h, C.byref(opts), C.c_void_p(x.data_ptr()), 0, 0))
_lib.check(L.zf_lasso_dev_stage(prob._h, 0))
n_slots = 120
evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n_slots)]
for k in range(n_slots):
    evs[k][0].record(stream)
    _lib.check(L.zf_lasso_dev_stage(prob._h, 1))
    evs[k][1].record(stream)
    _lib.check(L.zf_lasso_dev_stage(prob._h, 2))
    evs[k][2].record(stream)
torch.cuda.synchronize()
g_ms = np.array([e[0].elapsed_time(e[1]) for e in evs])[20:]
u_ms = np.array([e[1].elapsed_time(e[2]) for e in evs])[20:]
gap = np.array([evs[k][2].elapsed_time(evs[k + 1][0]) for k in range(n_slots - 1)])[20:]
tot = evs[20][0].elapsed_time(evs[-1][2]) / (n_slots - 20)
print(f"stages (events, {n_slots - 20} slots): gradient pass {g_ms.mean():.4f} ms (min {g_ms.min():.4f}), "
      f"update {1e3 * u_ms.mean():.1f} us (min {1e3 * u_ms.min():.1f}), gap {1e3 * gap.mean():.1f} us; "
      f"{tot:.4f} ms per slot")
