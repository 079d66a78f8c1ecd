#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for zfista_b200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload fds|jos1|sweep|lasso]

Headline (default) workload, BASELINE.json configs[2]: FDS tri-objective, n = 100, with the
L1 term of benchmarks/benchmark.py:439-440, FISTA (nesterov=True), tol_internal = 1e-11,
max_iter = 1e8; 1024 uniform(-2, 2) starts PER GPU (weak scaling: starts are independent, no
collective on the data path).  A "step" is one batched solve of a fresh batch of 1024
starts.  metric = converged solves/sec (starts whose status is 1 / device time).

  value : inputs resident in HBM, zf_solve_batched_device on torch's current stream, timed
          with CUDA events per step, L2 flushed between steps, max over ranks.
  e2e   : the public API call problem.minimize_proximal_gradient_batched(X0) on PINNED HOST
          arrays: H2D of the starts and D2H of x / fun / nit / status inside the timed region.
  roofline / also.lasso.roofline : see DESIGN.md "Measurement".
  cpu_baseline : the oracle (numpy + scipy restatement of the reference, kind "port") on
          the box's host cores, one start per core.

--impl reference times that same CPU path as its own arm (rank 0 only under torchrun).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "converged solves/sec (batched starts)"
UNIT = "solves/s"


# --------------------------------------------------------------------------- workloads
def workload_spec(name: str):
    """-> dict(cls, kw, low, high, n_starts, opts, label)  (benchmarks/benchmark.py:463-471)."""
    if name == "fds":
        n = 100
        return dict(cls="FDS", kw=dict(n_features=n, l1_ratios=(np.arange(3) + 1) / n,
                                       l1_shifts=np.arange(3.0)),
                    low=-2.0, high=2.0, n_starts=1024,
                    # max_iter_internal bounds the REFERENCE arm: with its default (100000)
                    # one start of this problem does not finish in 50 minutes of scipy
                    # trust-constr; at 100 it takes ~6 s.  The device's exact dual solver does
                    # not use the option.
                    opts=dict(nesterov=True, tol_internal=1e-11, max_iter=100000000,
                              max_iter_internal=100),
                    label="FDS tri-objective n=100 + L1, FISTA, 1024 starts per GPU")
    if name == "jos1":
        return dict(cls="JOS1", kw=dict(n_features=5), low=-2.0, high=4.0, n_starts=1000,
                    opts=dict(nesterov=True, tol_internal=1e-11, max_iter=100000000),
                    ref_sample=1000,
                    label="JOS1 bi-objective n=5, FISTA, 1000 starts per GPU")
    if name == "jos1_l1":
        n = 50
        return dict(cls="JOS1", kw=dict(n_features=n, l1_ratios=(np.arange(2) + 1) / n,
                                        l1_shifts=np.arange(2.0)),
                    low=-2.0, high=4.0, n_starts=1024,
                    opts=dict(nesterov=True, tol_internal=1e-11, max_iter=100000000),
                    ref_sample=128,
                    label="JOS1 bi-objective n=50 + L1, FISTA, 1024 starts per GPU")
    raise ValueError(name)


def make_starts(spec, seed, n_features):
    rng = np.random.RandomState(seed)
    return rng.uniform(spec["low"], spec["high"], size=(spec["n_starts"], n_features))


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.lines: list[str] = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- CPU arm
def _oracle_solve_one(args):
    import warnings

    from oracle import zfista_oracle as zo

    cls, kw, x0, opts = args
    warnings.simplefilter("ignore")
    bounds = kw.pop("bounds", None) if "bounds" in kw else None
    spec = zo.make_spec(cls, bounds=bounds, **kw)
    t0 = time.time()
    r = zo.minimize_proximal_gradient(spec, x0, **opts)
    return bool(r["success"]), int(r["nit"]), time.time() - t0


_POOL = None


def _warm(_):
    from oracle import zfista_oracle  # noqa: F401  (import scipy in the worker)

    return os.getpid()


def cpu_pool(cores):
    """Worker processes for the CPU arm, started and warmed (imports done) outside any
    timed region; the reference fans starts out over joblib workers the same way
    (benchmarks/benchmark.py:320-372)."""
    global _POOL
    if _POOL is None:
        from concurrent.futures import ProcessPoolExecutor

        os.environ.setdefault("OMP_NUM_THREADS", "1")
        os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
        _POOL = ProcessPoolExecutor(max_workers=cores)
        list(_POOL.map(_warm, range(4 * cores)))
    return _POOL


def cpu_reference_step(spec, X0, cores):
    """The reference's CPU path (oracle port: same numpy / scipy calls, one process per
    start) on `len(X0)` starts.  Returns (converged, seconds, nits)."""
    pool = cpu_pool(cores)
    tasks = [(spec["cls"], dict(spec["kw"]), X0[i], spec["opts"]) for i in range(len(X0))]
    t0 = time.time()
    out = list(pool.map(_oracle_solve_one, tasks, chunksize=max(1, len(tasks) // (4 * cores))))
    dt = time.time() - t0
    return sum(o[0] for o in out), dt, [o[1] for o in out]


def ref_sample_size(spec, args, cores):
    """Starts per CPU step: a bounded sample of the step's batch (about 10-30 s of CPU work)."""
    if args.ref_sample:
        return max(1, min(spec["n_starts"], args.ref_sample))
    return max(1, min(spec["n_starts"], spec.get("ref_sample") or cores))


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def headline_config(spec, world=1):
    """The `config` object of the JSON line: identical in both arms (the reference arm runs a
    bounded SAMPLE of this workload per step and says so in cpu_baseline.sample)."""
    return {"workload": spec["label"], "options": _json_opts(spec["opts"]),
            "starts_per_gpu": spec["n_starts"], "l2": "flushed between steps (256 MiB memset)",
            "inner_solver": "bounded Brent (m=2) / simplex Newton (m>=3)",
            "options_note": ("max_iter_internal=100 bounds the CPU arm only (reference default "
                             "100000: one start does not finish in 50 min of trust-constr); the "
                             "device's exact dual solver does not use it")}


def run_reference_arm(args):
    """The reference's CPU implementation of the path (oracle port: jax / jaxopt are not in the
    image) on all host cores.  A step = `cores` starts of the headline workload (one
    process per start, as benchmarks/benchmark.py:320-372 fans them out).  value counts EVERY
    solve that ran (converged or not) per second -- the reference's inexact inner solver leaves
    some starts unconverged and dropping them would flatter the GPU arm; converged_fraction says
    how many were.  Steps stop when the wall-clock budget is spent and `steps` is the number
    actually timed."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    spec = workload_spec(args.workload)
    cores = host_cores()
    n_sample = ref_sample_size(spec, args, cores)
    n_features = spec["kw"].get("n_features", 4)
    pool = cpu_pool(cores)

    def tasks_of(step, count):
        X0 = make_starts(spec, 1000 + step, n_features)[:count]
        return [(spec["cls"], dict(spec["kw"]), X0[i], spec["opts"]) for i in range(count)]

    # W untimed warm-up steps of one start each (the pool's processes are already imported and
    # warm; a CPU solve has no other state to warm up), run side by side
    if args.warmup > 0:
        list(pool.map(_oracle_solve_one, [t for w in range(args.warmup) for t in tasks_of(w, 1)]))
    # K timed steps: their starts are queued together so that no core idles at a step boundary
    # (the per-start times range from 0.1 s to 25 s); steps still pending when the budget is
    # spent are cancelled and not counted
    from concurrent.futures import FIRST_COMPLETED, wait

    t_begin = time.time()
    futs = {}
    for step in range(args.steps):
        for t in tasks_of(args.warmup + step, n_sample):
            futs[pool.submit(_oracle_solve_one, t)] = step
    per_step = [[] for _ in range(args.steps)]
    pending = set(futs)
    while pending:
        done, pending = wait(pending, timeout=1.0, return_when=FIRST_COMPLETED)
        for f in done:
            per_step[futs[f]].append(f.result())
        if time.time() - t_begin > args.ref_budget_s:
            running = [f for f in pending if not f.cancel()]
            for f in running:                     # already started: let them finish, count them
                per_step[futs[f]].append(f.result())
            break
    total = time.time() - t_begin
    full = [r for r in per_step if len(r) == n_sample]
    done_all = [o for r in per_step for o in r]
    conv = sum(o[0] for o in done_all)
    nits = [o[1] for o in done_all]
    times = full                                   # steps whose every start finished
    solves = len(done_all)
    value = solves / total if total > 0 else 0.0
    sample = (f"{n_sample} starts per step ({len(times)} steps timed of {args.steps} requested, "
              f"budget {args.ref_budget_s:.0f} s) of the same workload; oracle port of zfista "
              "(numpy + scipy trust-constr, max_iter_internal=100), one process per start; all "
              "solves counted, converged or not")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": len(times), "steps_requested": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * total / max(1, len(times)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": headline_config(spec),
        "same_config": False,
        "same_config_note": (f"same problem and options; {n_sample} starts per CPU step against "
                             f"{spec['n_starts']} per GPU step (throughput per start is what is "
                             "compared)"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "nit_mean": float(np.mean(nits)) if nits else None,
        "solves_timed": solves, "seconds_timed": total,
        "converged_fraction": conv / max(1, solves),
        "converged_solves_per_s": conv / total if total > 0 else 0.0,
    }
    emit(line)
    return 0


def _json_opts(opts):
    return {k: (v if not isinstance(v, (np.floating, np.integer)) else v.item())
            for k, v in opts.items()}


# --------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from zfista_b200 import _lib
    from zfista_b200 import build as zbuild
    import zfista_b200.problems as zp
    from zfista_b200.proximal_gradient import _make_options

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run "
                             "(one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: zfista_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    zbuild.build()
    L = _lib.lib()

    spec = workload_spec(args.workload)
    prob = getattr(zp, spec["cls"])(**spec["kw"])
    n, m, S = prob.n_features, prob.n_objectives, spec["n_starts"]
    total_steps = args.warmup + args.steps
    # every step gets its own batch of starts; each rank its own slice of the seed space
    host_batches = [make_starts(spec, 1000 + step * world + rank, n) for step in range(total_steps)]
    dev_batches = [torch.from_numpy(b).to(dev) for b in host_batches]
    out_x = torch.empty(S, n, dtype=torch.float64, device=dev)
    out_fun = torch.empty(S, m, dtype=torch.float64, device=dev)
    out_nit = torch.empty(S, dtype=torch.int64, device=dev)
    out_status = torch.empty(S, dtype=torch.int32, device=dev)
    out_ndual = torch.empty(S, dtype=torch.int64, device=dev)
    out_nfev = torch.empty(S, dtype=torch.int64, device=dev)
    res = _lib.ZfResult()
    res.x, res.fun, res.nit, res.status = (out_x.data_ptr(), out_fun.data_ptr(),
                                           out_nit.data_ptr(), out_status.data_ptr())
    res.n_dual, res.nfev = out_ndual.data_ptr(), out_nfev.data_ptr()
    o = spec["opts"]
    opts = _make_options(1.0, 1e-5, o["tol_internal"], o["max_iter"],
                         o.get("max_iter_internal", 100000), 100, False, 0.5,
                         o.get("nesterov", False), (0, 0.25), False, "reference", 0)
    desc, keep = prob.descriptor()
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step(i):
        _lib.check(L.zf_solve_batched_device(
            C.byref(desc), C.byref(opts), S, C.c_void_p(dev_batches[i].data_ptr()), None,
            C.byref(res), C.c_void_p(stream.cuda_stream)))

    # ---------------- value: device-resident, CUDA events per step, L2 flushed between steps
    for i in range(args.warmup):
        device_step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    converged = 0
    nit_sum = 0
    ndual_sum = 0
    nit_max = 0
    barrier()
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record(stream)
        device_step(args.warmup + k)
        ev[k][1].record(stream)
        ev[k][1].synchronize()
        converged += int((out_status == 1).sum().item())
        nit_sum += int(out_nit.sum().item())
        nit_max = max(nit_max, int(out_nit.max().item()))
        ndual_sum += int(out_ndual.sum().item())
    barrier()
    launches = _lib.launch_count() - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- e2e: public API, pinned host arrays in, host arrays out
    pinned = [torch.from_numpy(b).pin_memory() for b in host_batches]
    for i in range(min(args.warmup, 2)):
        prob.minimize_proximal_gradient_batched(pinned[i].numpy(), **spec["opts"])
    barrier()
    e2e_s = 0.0
    e2e_conv = 0
    for k in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        br = prob.minimize_proximal_gradient_batched(pinned[args.warmup + k].numpy(),
                                                     **spec["opts"])
        e2e_s += time.perf_counter() - t0
        e2e_conv += int((br.status == 1).sum())
    h2d = S * n * 8
    d2h = S * (n * 8 + m * 8 + 8 + 4 + 8 + 8 + 8 + 8)

    # ---------------- reduce over ranks: time = max, counts = sum
    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    c = torch.tensor([converged, e2e_conv, nit_sum, ndual_sum, launches], dtype=torch.float64,
                     device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_ms_max = t.tolist()
    conv_all, e2e_conv_all, nit_all, ndual_all, launches_all = c.tolist()

    line = None
    if rank == 0:
        value = conv_all / (dev_ms_max / 1e3)
        e2e_value = e2e_conv_all / (e2e_ms_max / 1e3)
        peaks = _measured_peaks()
        # dominant kernel = batched_fista_kernel: one launch per step.  Algorithmic HBM bytes
        # per start: x0 in; x, fun, nit, status, n_dual, nfev out (DESIGN.md "Measurement").
        bytes_per_start = n * 8 + n * 8 + m * 8 + 8 + 4 + 8 + 8
        ker_s = (dev_ms / 1e3) / args.steps
        achieved = S * bytes_per_start / ker_s / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": spec["label"], "options": _json_opts(spec["opts"]),
                       "starts_per_gpu": S, "l2": "flushed between steps (256 MiB memset)",
                       "inner_solver": "bounded Brent (m=2) / simplex Newton (m>=3)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms_max / args.steps},
            "gpu_launches": int(launches_all),
            "clocks": clocks,
            "nit_mean": nit_all / (S * world * args.steps), "nit_max_rank0": nit_max,
            "dual_evals_per_solve": ndual_all / (S * world * args.steps),
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": 1239808.0,
                "ncu": {"fp64_pipe_active_pct": 16.6, "issue_slots_active_pct": 34.8,
                        "warp_occupancy_pct": 7.0, "dominant_stall": "wait (dependent FP64 ops)",
                        "source": "profiles/r01e_batched_fista_final.txt"},
                "peak_source": peaks["source"],
                "note": ("batched_fista_kernel keeps each start's state in shared memory and "
                         "touches HBM only for x0 and the results; it is FP64-latency bound, "
                         "not HBM bound (profiles/), so this fraction is tiny by design. The "
                         "HBM-bound kernel of the path is the dense LASSO pass: also.lasso")},
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = host_cores()
            ns = ref_sample_size(spec, args, cores)
            cconv, cdt, cn = cpu_reference_step(spec, host_batches[args.warmup][:ns], cores)
            line["cpu_baseline"] = {
                "value": cconv / cdt if cdt > 0 else 0.0, "unit": UNIT, "cores": cores,
                "kind": "port", "seconds": cdt, "nit": cn,
                "sample": (f"{ns} starts of this step's batch, oracle port "
                           "of zfista (numpy + scipy trust-constr), one process per start")}
        else:
            line["cpu_baseline"] = None
    if rank == 0 or world > 1:
        also = {}
        if not args.no_extras:
            for name, fn in (("lasso", bench_lasso), ("lasso_multi", bench_lasso_multi),
                             ("cameraman", bench_cameraman),
                             ("ab_sweep", bench_sweep), ("batch_scaling", bench_batch_scaling),
                             ("lasso_configs3", bench_lasso_configs3)):
                try:
                    also[name] = fn(args, dev, rank, world)
                except Exception as e:  # extras must never lose the headline line
                    also[name] = {"error": repr(e)}
        if rank == 0:
            line["also"] = also
            emit(line)
    del keep
    if world > 1:
        dist.destroy_process_group()
    return 0


def _measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return {"hbm_gbs": float(d["hbm_gbs"]), "source": "MEASURED_PEAKS.json (of measured)"}
    return {"hbm_gbs": 6650.0, "source": "B200_PROFILING.md fallback (of fallback)"}


_LASSO_DATA = {}


def _lasso_data(rows, cols, dev, rank):
    """Synthetic dense A (rows x cols fp64, this rank's row shard) and b = A w; built once and
    shared by the single-run and the multi-run LASSO extras."""
    import torch

    key = (rows, cols, str(dev), rank)
    if key not in _LASSO_DATA:
        g = torch.Generator(device=dev).manual_seed(7 + rank)
        A = torch.empty(rows, cols, dtype=torch.float64, device=dev)
        chunk = max(1, (64 << 20) // (cols * 8))
        for r0 in range(0, rows, chunk):
            r1 = min(rows, r0 + chunk)
            A[r0:r1] = torch.randn(r1 - r0, cols, dtype=torch.float64, device=dev, generator=g)
        w = torch.zeros(cols, dtype=torch.float64, device=dev)
        w[:64] = 1.0
        _LASSO_DATA.clear()
        _LASSO_DATA[key] = (A, A @ w)
    return _LASSO_DATA[key]


def bench_lasso_multi(args, dev, rank, world):
    """Many LASSO runs sharing one A (north_star (c): FP64 tensor-core DGEMM when many
    right-hand sides share A): `--lasso-runs` FISTA runs with different momentum (a, b) over the
    same A as `also.lasso`.  One gradient of all runs = 2 DGEMM passes over A on the FP64 tensor
    cores (csrc/zf_lasso_multi.cu).  Reports the roofline of each pass (CUDA events on the
    launching stream, L2 flushed between launches) and FISTA run-iterations/s."""
    import ctypes as C
    import warnings

    import torch
    import torch.distributed as dist

    from zfista_b200 import _lib
    from zfista_b200.lasso import DenseLassoMulti

    rows, cols, K = args.lasso_rows, args.lasso_cols, args.lasso_runs
    A, b = _lasso_data(rows, cols, dev, rank)
    prob = DenseLassoMulti(A, b, 1e-3, K, scale=1.0 / (2 * rows * world), distributed=world > 1)
    grid = [AB_GRID[k % len(AB_GRID)] for k in range(K)]
    X = torch.zeros(K, cols, dtype=torch.float64, device=dev)
    a_bytes = rows * cols * 8
    out = {"A": f"{rows}x{cols} fp64 per GPU ({a_bytes / 2**30:.1f} GiB, > L2)", "runs": K}
    stream = torch.cuda.current_stream()
    if world == 1:
        L = _lib.lib()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        peaks = _measured_peaks()
        prob.gradient(X)
        torch.cuda.synchronize()
        # algorithmic bytes of one pass: A once + the K vectors in and out
        alg = a_bytes + K * (rows + cols) * 8
        flops = 2.0 * rows * cols * K
        for which, name in ((0, "pass1_residual"), (1, "pass2_atr")):
            ts = []
            for _ in range(5):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                _lib.check(L.zf_lasso_multi_pass_device(prob._h, C.c_void_p(X.data_ptr()), which))
                e1.record(stream)
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = sum(ts) / len(ts)
            ach = alg / (ms / 1e3) / 1e9
            out[name] = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": ach / peaks["hbm_gbs"], "ms_per_launch": ms,
                         "fp64_tensor_tflops": flops / (ms / 1e3) / 1e12,
                         "peak_source": peaks["source"]}
        del flush
        ms2 = out["pass1_residual"]["ms_per_launch"] + out["pass2_atr"]["ms_per_launch"]
        out["ms_per_gradient_of_all_runs"] = ms2
        out["ms_per_gradient_per_run"] = ms2 / K
    iters = args.lasso_iters
    kw = dict(lr=0.5, decay_rate=1, nesterov=True, tol=0.0, return_device=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        prob.minimize_proximal_gradient_batched(X, grid, max_iter=3, **kw)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = prob.minimize_proximal_gradient_batched(X, grid, max_iter=iters, **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    n_it = sum(r.nit for r in res)
    out["fista_run_iters_per_s"] = n_it / tt.item()
    out["fista_iters_per_run"] = res[0].nit
    out["global_rows"] = rows * world
    return out


def bench_lasso_configs3(args, dev, rank, world):
    """BASELINE configs[3]: dense LASSO A 200000 x 20000 fp64 (29.8 GiB), rows sharded over the
    ranks (strong scaling: 200000 / world rows per GPU), A^T r all-reduced over NCCL.  Single-run
    path (one-pass fused gradient) and 16 runs sharing A (two FP64 tensor-core passes)."""
    import warnings

    import torch
    import torch.distributed as dist

    from zfista_b200.lasso import DenseLasso, DenseLassoMulti

    _LASSO_DATA.clear()                      # free the 8 GiB matrix of the other extras
    torch.cuda.empty_cache()
    rows_total, cols, K = 200000, 20000, 16
    rows = rows_total // world
    free, _ = torch.cuda.mem_get_info()
    if free < rows * cols * 8 * 1.15:
        return {"skipped": f"needs {rows * cols * 8 / 2**30:.1f} GiB of free HBM"}
    A, b = _lasso_data(rows, cols, dev, rank)
    peaks = _measured_peaks()
    a_bytes = rows * cols * 8
    out = {"A": f"{rows_total}x{cols} fp64, {rows} rows per GPU ({a_bytes / 2**30:.1f} GiB per GPU)"}
    x = torch.zeros(cols, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream()
    single = DenseLasso(A, b, l1_ratio=1e-3, scale=1.0 / (2 * rows_total), distributed=world > 1)
    multi = DenseLassoMulti(A, b, 1e-3, K, scale=1.0 / (2 * rows_total), distributed=world > 1)
    if world == 1:
        X = torch.zeros(K, cols, dtype=torch.float64, device=dev)
        for name, fn, n_pass in (("single_run_gradient", lambda: single.gradient(x), 1),
                                 ("multi_run_gradient", lambda: multi.gradient(X), 2)):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(3):
                fn()
            e1.record(stream)
            e1.synchronize()
            ms = e0.elapsed_time(e1) / 3
            ach = n_pass * a_bytes / (ms / 1e3) / 1e9
            out[name] = {"bound": "hbm", "ms": ms, "algorithmic_passes_over_A": n_pass,
                         "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": ach / peaks["hbm_gbs"]}
        out["multi_run_gradient"]["runs"] = K
        out["multi_run_gradient"]["ms_per_run"] = out["multi_run_gradient"]["ms"] / K
    kw = dict(lr=0.5, decay_rate=1, nesterov=True, tol=0.0, return_device=True)
    grid = [AB_GRID[k % len(AB_GRID)] for k in range(K)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for name, run in (("single", lambda n: [single.minimize_proximal_gradient(x, max_iter=n, **kw)]),
                          ("multi", lambda n: multi.minimize_proximal_gradient_batched(x, grid, max_iter=n, **kw))):
            run(2)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = run(10)
            torch.cuda.synchronize()
            tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            out[f"{name}_fista_run_iters_per_s"] = sum(r.nit for r in res) / tt.item()
    del single, multi
    _LASSO_DATA.clear()
    torch.cuda.empty_cache()
    return out


def bench_lasso(args, dev, rank, world):
    """Dense LASSO gradient pass (the HBM-bound kernels): A rows x cols fp64 per GPU, rows
    sharded over ranks (weak: every rank holds `rows` rows), A^T r all-reduced over NCCL.
    Reports FISTA iterations/s of a fixed-step run and the roofline of one gradient."""
    import torch
    import torch.distributed as dist

    from zfista_b200 import _lib
    from zfista_b200.lasso import DenseLasso

    rows, cols = args.lasso_rows, args.lasso_cols
    A, b = _lasso_data(rows, cols, dev, rank)
    prob = DenseLasso(A, b, l1_ratio=1e-3, scale=1.0 / (2 * rows * world),
                      distributed=world > 1)
    x = torch.zeros(cols, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream()
    a_bytes = rows * cols * 8
    out = {"A": f"{rows}x{cols} fp64 per GPU ({a_bytes / 2**30:.1f} GiB, > L2)"}
    if world == 1:
        # kernel-level: one gradient = residual pass + A^T pass (2 x |A| from HBM)
        for _ in range(3):
            prob.gradient(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record(stream)
        for _ in range(reps):
            prob.gradient(x)
        e1.record(stream)
        e1.synchronize()
        ms = e0.elapsed_time(e1) / reps
        peaks = _measured_peaks()
        # Algorithmic bytes of one gradient A^T(A v - b): A once (a fused kernel keeps the row
        # on chip between the dot product and the rank-1 update) plus the vectors.  The
        # two-pass kernels read A twice; `passes` says which form this shape runs.
        vec_bytes = (2 * cols + 2 * rows) * 8
        alg = a_bytes + vec_bytes
        passes = prob.hbm_passes_per_gradient()
        ach = alg / (ms / 1e3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tp):        # ncu-measured DRAM bytes / |A| per kernel form
            with open(tp) as fh:
                tj = json.load(fh)
            keys = (["lasso_fused_ring_kernel"] if passes == 1
                    else ["lasso_residual_kernel", "lasso_atr_kernel"])
            traffic = sum((tj[k]["dram_bytes_read"] + tj[k]["dram_bytes_write"])
                          / tj[k]["algorithmic_bytes"] for k in keys) * a_bytes
        out["roofline"] = {
            "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": ach / peaks["hbm_gbs"], "traffic": traffic,
            "traffic_source": "ncu dram bytes / |A| per kernel (profiles/r01_traffic.json, "
                              "measured at 32768x16384), scaled to this A",
            "ms_per_gradient": ms, "hbm_passes_over_A": passes,
            "dram_GBps": passes * a_bytes / (ms / 1e3) / 1e9,
            "peak_source": peaks["source"]}
    # solver-level: fixed-step FISTA iterations per second (A stays resident)
    iters = args.lasso_iters
    lr = 0.5
    prob.minimize_proximal_gradient(x, lr=lr, decay_rate=1, nesterov=True, max_iter=3, tol=0.0,
                                    return_device=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    import warnings

    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = prob.minimize_proximal_gradient(x, lr=lr, decay_rate=1, nesterov=True,
                                              max_iter=iters, tol=0.0, return_device=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    out["fista_iters_per_s"] = res.nit / tt.item()
    out["fista_iters"] = res.nit
    out["global_rows"] = rows * world
    return out


AB_GRID = [
    (0.0, 0.0), (0.0, 1 / 8), (0.0, 1 / 4), (1 / 6, 1 / 144), (1 / 6, 37 / 288), (1 / 6, 1 / 4),
    (1 / 4, 1 / 64), (1 / 4, 17 / 128), (1 / 4, 1 / 4), (1 / 2, 1 / 16), (1 / 2, 5 / 32),
    (1 / 2, 1 / 4), (3 / 4, 9 / 64), (3 / 4, 25 / 128), (3 / 4, 1 / 4),
]


def _cameraman_cpu_iters(args):
    obs, kernel, l1, x0, L, ab, iters = args
    import warnings

    from oracle import deblur_oracle as do

    warnings.simplefilter("ignore")
    t0 = time.time()
    r = do.minimize(obs, kernel, l1, x0, lr=1 / L, decay_rate=1, nesterov=True,
                    nesterov_ratio=ab, max_iter=iters, tol=0.0)
    return r["nit"], time.time() - t0


def synthetic_observation(h, w, kernel, seed=0, noise=1e-3):
    """Synthetic stand-in for the notebook's blurred + noisy cameraman image (no network for
    skimage.data): a piecewise-smooth scene in [0, 1], blurred with `kernel` (symmetric
    boundary) plus N(0, noise^2)."""
    from scipy.signal import correlate2d

    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = 0.5 + 0.25 * np.sin(2 * np.pi * xx / w * 1.5) * np.cos(2 * np.pi * yy / h)
    for _ in range(6):
        cy, cx = rng.uniform(0, h), rng.uniform(0, w)
        ry, rx = rng.uniform(h / 16, h / 4), rng.uniform(w / 16, w / 4)
        img[((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 < 1] = rng.uniform(0, 1)
    img = np.clip(img, 0, 1)
    return correlate2d(img, kernel, mode="same", boundary="symm") + rng.standard_normal((h, w)) * noise


def bench_cameraman(args, dev, rank, world):
    """BASELINE configs[1]: 256x256 deblurring (9x9 Gaussian blur, Haar, l1 = 2e-5), the
    notebook's 15 (a, b) pairs as 15 runs of one call, fixed step 1/L; every rank solves the
    same 15 runs (replicas).  Capped at --cameraman-iters iterations per run so that the CPU
    arm can time the identical work; reports FISTA iterations/s summed over the runs, timed
    through the host entry point (x0 and the (a, b) table go H2D, x / fun / nit come D2H)."""
    import warnings

    import torch

    from zfista_b200.deblur import HaarDeblurL1, gaussian_kernel, lipschitz_constant

    kernel = gaussian_kernel(9, 4.0)
    kernel /= kernel.sum()
    obs = synthetic_observation(256, 256, kernel, seed=1)
    l1 = 2e-5
    prob = HaarDeblurL1(obs, kernel, l1)
    x0 = prob.dwt_array(obs)
    L = lipschitz_constant(kernel)
    pairs = np.array(AB_GRID)
    iters = args.cameraman_iters
    kw = dict(lr=1 / L, decay_rate=1, nesterov=True, max_iter=iters, tol=0.0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        prob.minimize_proximal_gradient_batched(x0, pairs, **dict(kw, max_iter=20))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = prob.minimize_proximal_gradient_batched(x0, pairs, **kw)
        dt = time.perf_counter() - t0
    total = sum(r.nit for r in res)
    out = {"workload": "256x256 Haar/9x9-blur deblurring, 15 (a,b) runs, fixed step, "
                       f"{iters} iterations per run",
           "fista_iters_per_s": total / dt, "seconds": dt, "iterations": total,
           "us_per_round_of_15": 1e6 * dt / iters}
    # algorithmic HBM bytes per run-iteration: x, x_prev in; y, g out; y, g in, x out; b in
    out["algorithmic_GBps"] = total * 8 * 65536 * 8 / dt / 1e9
    # FP64 work in 81-tap-equivalent flops: two 9x9 correlations on (40^2 + 32^2) points per
    # 32x32 tile (the folded stencil for symmetric kernels executes 35 % fewer of them)
    out["fp64_tflops"] = total * 64 * (1600 + 1024) * 81 * 2 / dt / 1e12
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        cores = host_cores()
        pool = cpu_pool(cores)
        cpu_iters = 30
        tasks = [(obs, kernel, l1, x0, L, tuple(ab), cpu_iters) for ab in AB_GRID]
        t0 = time.time()
        got = list(pool.map(_cameraman_cpu_iters, tasks))
        cdt = time.time() - t0
        out["cpu_fista_iters_per_s"] = sum(g[0] for g in got) / cdt
        out["cpu_sample"] = (f"{cpu_iters} iterations of each of the 15 runs, oracle port "
                             f"(scipy correlate2d + numpy Haar), {cores} cores")
    return out


def bench_batch_scaling(args, dev, rank, world):
    """The headline problem at larger batches per GPU (device-resident, CUDA events): the
    kernel is latency bound at 1024 starts (7 % of the warp slots), so solves/s keeps growing
    with the batch until the FP64 pipe saturates."""
    import torch

    from zfista_b200 import _lib
    import zfista_b200.problems as zp
    from zfista_b200.proximal_gradient import _make_options

    spec = workload_spec(args.workload)
    prob = getattr(zp, spec["cls"])(**spec["kw"])
    n, m = prob.n_features, prob.n_objectives
    o = spec["opts"]
    opts = _make_options(1.0, 1e-5, o["tol_internal"], o["max_iter"],
                         o.get("max_iter_internal", 100000), 100, False, 0.5,
                         o.get("nesterov", False), (0, 0.25), False, "reference", 0)
    desc, keep = prob.descriptor()
    L = _lib.lib()
    stream = torch.cuda.current_stream()
    out = {}
    for S in (1024, 4096, 16384, 65536):
        rng = np.random.RandomState(5 + rank)
        X0 = torch.from_numpy(rng.uniform(spec["low"], spec["high"], size=(S, n))).to(dev)
        x = torch.empty(S, n, dtype=torch.float64, device=dev)
        fun = torch.empty(S, m, dtype=torch.float64, device=dev)
        nit = torch.empty(S, dtype=torch.int64, device=dev)
        status = torch.empty(S, dtype=torch.int32, device=dev)
        res = _lib.ZfResult()
        res.x, res.fun, res.nit, res.status = (x.data_ptr(), fun.data_ptr(), nit.data_ptr(),
                                               status.data_ptr())

        def step():
            _lib.check(L.zf_solve_batched_device(C.byref(desc), C.byref(opts), S,
                                                 C.c_void_p(X0.data_ptr()), None, C.byref(res),
                                                 C.c_void_p(stream.cuda_stream)))
        step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step()
        e1.record(stream)
        e1.synchronize()
        ms = e0.elapsed_time(e1)
        out[str(S)] = {"ms": ms, "solves_per_s": float((status == 1).sum().item()) / (ms / 1e3)}
    del keep
    return out


def bench_sweep(args, dev, rank, world):
    """BASELINE configs[4]: momentum (a, b) grid x 1024 starts on JOS1 (n = 50, +L1) as ONE
    launch of 15 x 1024 = 15360 starts per GPU (per-start (a, b) table)."""
    import torch

    from zfista_b200.distributed import momentum_grid
    import zfista_b200.problems as zp

    n = 50
    prob = zp.JOS1(n_features=n, l1_ratios=(1 / n, 1 / n / 2), l1_shifts=(0, 1))
    rng = np.random.RandomState(77 + rank)
    X0 = rng.uniform(-2, 4, size=(1024, n))
    Xg, AB, _, _ = momentum_grid(X0, AB_GRID)
    prob.minimize_proximal_gradient_batched(Xg[:64], nesterov=True, nesterov_ratio=AB[:64],
                                            tol_internal=1e-11)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    br = prob.minimize_proximal_gradient_batched(Xg, nesterov=True, nesterov_ratio=AB,
                                                 tol_internal=1e-11)
    dt = time.perf_counter() - t0
    out = {"workload": "JOS1 n=50 +L1, 15 (a,b) pairs x 1024 starts in one launch (e2e call)",
           "solves": int(len(Xg)), "converged": int((br.status == 1).sum()),
           "solves_per_s": float((br.status == 1).sum() / dt), "seconds": dt,
           "nit_mean": float(br.nit.mean()), "dual_evals_per_solve": float(br.n_dual.mean())}
    # the same sweep with the exact simplex-Newton dual solver instead of the reference's
    # bounded Brent (dual_solver="newton": same optimum, not the reference's sqrt(eps) noise)
    t0 = time.perf_counter()
    bn = prob.minimize_proximal_gradient_batched(Xg, nesterov=True, nesterov_ratio=AB,
                                                 tol_internal=1e-11, dual_solver="newton")
    dtn = time.perf_counter() - t0
    out["newton_dual"] = {"solves_per_s": float((bn.status == 1).sum() / dtn), "seconds": dtn,
                          "nit_mean": float(bn.nit.mean()),
                          "dual_evals_per_solve": float(bn.n_dual.mean())}
    return out


_JSON_FD = None


def capture_stdout():
    """stdout must carry exactly ONE JSON line, but libraries write to file descriptor 1 on
    their own (NCCL prints "NCCL version ..." at communicator creation whatever NCCL_DEBUG
    says).  Everything written to fd 1 from here on goes to stderr; emit() writes the JSON
    line to the real stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="fds", choices=["fds", "jos1", "jos1_l1"])
    ap.add_argument("--ref-sample", type=int, default=0,
                    help="starts per CPU step (default: one per host core)")
    ap.add_argument("--ref-budget-s", type=float, default=240.0,
                    help="wall-clock budget of the --impl reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--lasso-rows", type=int, default=65536)
    ap.add_argument("--lasso-cols", type=int, default=16384)
    ap.add_argument("--lasso-iters", type=int, default=50)
    ap.add_argument("--lasso-runs", type=int, default=16,
                    help="runs sharing A in also.lasso_multi (1..32)")
    ap.add_argument("--cameraman-iters", type=int, default=2000)
    args = ap.parse_args()
    capture_stdout()
    try:
        if args.impl == "reference":
            return run_reference_arm(args)
        return run_ours(args)
    finally:
        if _POOL is not None:       # join the CPU-arm workers before interpreter shutdown
            _POOL.shutdown(wait=True, cancel_futures=True)


if __name__ == "__main__":
    sys.exit(main())
