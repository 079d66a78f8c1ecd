"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` for the control plane.

Two shapes of parallelism exist on the path (BASELINE north_star):

* **independent starts / (a, b) grid points** shard over ranks with NO data-path collective:
  every rank solves its contiguous slice of the batch on its own GPU
  (:func:`minimize_proximal_gradient_sharded`).  Results can be all-gathered afterwards
  (control plane, off the timed path).
* **the dense LASSO shards A by rows**; its one exchange per pass is the all-reduce of
  ``[A^T r | sum r^2]`` (:func:`run_split_lasso`, used by ``DenseLasso`` with NCCL).

Both functions take the local compute as a parameter so that the sharding / exchange logic
is testable on CPU with the ``gloo`` backend (tests/test_distributed_cpu.py); the defaults
are the CUDA paths.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous balanced partition: the first ``n_items % world`` ranks get one extra."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("need 0 <= rank < world")
    base, extra = divmod(int(n_items), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def momentum_grid(X0, grid):
    """Cartesian product starts x (a, b) pairs as one batch (the
    PGM_experiment_with_various_a_b workload): returns (X0_rep, AB, start_index, grid_index)
    with row ``s * len(grid) + g`` = start s under pair g."""
    X0 = np.asarray(X0, dtype=np.float64)
    grid = np.asarray(grid, dtype=np.float64).reshape(-1, 2)
    S, G = len(X0), len(grid)
    return (np.repeat(X0, G, axis=0), np.tile(grid, (S, 1)),
            np.repeat(np.arange(S), G), np.tile(np.arange(G), S))


def _dist(group):
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        return None, 0, 1
    return dist, dist.get_rank(group), dist.get_world_size(group)


def minimize_proximal_gradient_sharded(problem, X0, nesterov_ratio=(0, 0.25), group=None,
                                       gather=True, local_solver=None, **kwargs):
    """Solve every row of ``X0`` with the rows sharded over the ranks of ``group``.

    Every rank passes the SAME ``X0`` (and per-start ``nesterov_ratio`` table, if any) and
    solves rows ``shard_bounds(len(X0), rank, world)``.  With ``gather=True`` every rank
    returns the full :class:`~zfista_b200.proximal_gradient.BatchResult` in the original
    row order; with ``gather=False`` only its own slice (no communication at all)."""
    from .proximal_gradient import BatchResult, minimize_proximal_gradient_batched

    dist, rank, world = _dist(group)
    X0 = np.ascontiguousarray(np.asarray(X0, dtype=np.float64))
    lo, hi = shard_bounds(len(X0), rank, world)
    ab = np.asarray(nesterov_ratio, dtype=np.float64)
    local_ab = ab[lo:hi] if ab.ndim == 2 else nesterov_ratio
    solve = local_solver or minimize_proximal_gradient_batched
    local = solve(problem, X0[lo:hi], nesterov_ratio=local_ab, **kwargs)
    if not gather or world == 1:
        return local
    fields = ("x", "fun", "nit", "status", "lr", "nfev", "n_dual", "err")
    payload = {k: getattr(local, k) for k in fields}
    payload["time"] = local.time
    parts = [None] * world
    dist.all_gather_object(parts, payload, group=group)
    merged = {k: np.concatenate([p[k] for p in parts], axis=0) for k in fields}
    return BatchResult(time=max(p["time"] for p in parts), **merged)


def run_split_lasso(ops, allreduce):
    """Drive one solve through the split LASSO protocol of include/zfista_b200.h:

        begin -> [allreduce(partial)] -> step -> { grad(next) -> [allreduce] -> step }* -> finish

    ``ops`` provides ``begin()``, ``grad(which)``, ``step() -> next`` (0: gradient at y,
    1: f at the candidate, 2: done) and ``finish()``; ``allreduce()`` sums the partial
    buffer over the row shards (a no-op on one GPU).  Every rank takes identical decisions
    because every rank sees the identical reduced buffer."""
    ops.begin()
    allreduce()
    nxt = ops.step()
    while nxt != 2:
        ops.grad(nxt)
        allreduce()
        nxt = ops.step()
    return ops.finish()


def run_device_lasso(ops, allreduce_all, allreduce_ss, chunk=16):
    """Drive one row-sharded solve through the DEVICE-decided protocol of
    include/zfista_b200.h (zf_lasso_dev_*): every decision is taken on the GPU, the host only
    enqueues trials -- stage by stage, with the all-reduce of ``partial`` between the stages, all
    on one stream -- and polls a flag one chunk of trials behind the one it is enqueuing.

    ``ops``: ``begin()``, ``stage(k)``, ``needs_feval() -> bool``, ``snapshot(slot)``,
    ``wait(slot) -> done``, ``finish()``.  ``allreduce_all()`` sums ``[A^T r | sum r^2]`` over the
    row shards, ``allreduce_ss()`` only the last value (no-ops on one GPU).  Every rank enqueues
    the same sequence and sees the same reduced values, so every rank's GPU takes the same
    decisions."""
    ops.begin()
    allreduce_ss()
    ops.stage(0)
    feval = ops.needs_feval()

    def enqueue_chunk():
        for _ in range(chunk):
            ops.stage(1)
            allreduce_all()
            ops.stage(2)
            if feval:
                ops.stage(3)
                allreduce_ss()
                ops.stage(4)

    cur = 0
    enqueue_chunk()
    ops.snapshot(cur)
    while True:
        enqueue_chunk()
        ops.snapshot(1 - cur)
        if ops.wait(cur):
            break
        cur = 1 - cur
    if not feval:
        ops.stage(6)
        allreduce_ss()
    ops.stage(5)
    return ops.finish()
