"""Where the end-to-end call of the headline spends its time beyond the kernel: Python around the
C call, the C call, the kernel (device-resident call timed with CUDA events).

    python profiles/time_e2e_breakdown.py"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from zfista_b200 import _lib, proximal_gradient as pg  # noqa: E402

spec = bench.workload_spec("fds")
runner = bench.BatchedRunner(spec, torch.device("cuda", 0), 0, 1, 12)
prob, opts = runner.prob, spec["opts"]
pinned = [torch.from_numpy(b).pin_memory() for b in runner.host_batches]
L = _lib.lib()
real = L.zf_solve_batched_host
acc = {"c": 0.0}


def timed(*a):
    t0 = time.perf_counter()
    rc = real(*a)
    acc["c"] += time.perf_counter() - t0
    return rc


class _Wrap:
    def __getattr__(self, name):
        return timed if name == "zf_solve_batched_host" else getattr(L, name)


for i in range(2):
    prob.minimize_proximal_gradient_batched(pinned[i].numpy(), **opts)
_orig = _lib.lib
_lib.lib = lambda: _Wrap()
tot = []
for k in range(2, 12):
    torch.cuda.synchronize()
    acc["c"] = 0.0
    t0 = time.perf_counter()
    prob.minimize_proximal_gradient_batched(pinned[k].numpy(), **opts)
    tot.append((time.perf_counter() - t0, acc["c"]))
_lib.lib = _orig
kern = []
for k in range(2, 12):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(runner.stream)
    runner.device_step(k)
    e1.record(runner.stream)
    torch.cuda.synchronize()
    kern.append(e0.elapsed_time(e1))
t = np.array(tot) * 1e3
kern = np.array(kern)
print(f"same 10 batches, medians: e2e call {np.median(t[:, 0]):.3f} ms | C call {np.median(t[:, 1]):.3f} ms | "
      f"python around it {np.median(t[:, 0] - t[:, 1]):.3f} ms | device-resident step "
      f"{np.median(kern):.3f} ms | C call minus device step {np.median(t[:, 1] - kern):.3f} ms")
