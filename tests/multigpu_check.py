"""Run under torchrun on N >= 2 GPUs (tests/test_gpu_multi.py launches it):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multigpu_check.py

Checks, with NCCL: (1) the row-sharded DenseLasso gives the single-GPU result and the CPU
oracle's; (2) starts sharded over ranks (no collective) give the single-GPU batch result;
(3) the row-sharded shared-A multi-run path (DenseLassoMulti, [A^T R | sum r^2] of every run
all-reduced) gives, run by run, the oracle's result."""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from oracle import zfista_oracle as zo
    from zfista_b200 import distributed as zd
    from zfista_b200.lasso import DenseLasso, DenseLassoMulti
    import zfista_b200.problems as zp

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    rng = np.random.RandomState(11)
    n_rows, n_cols = 403, 130
    A = rng.standard_normal((n_rows, n_cols))
    w = np.zeros(n_cols)
    w[:8] = rng.standard_normal(8)
    b = A @ w + 0.01 * rng.standard_normal(n_rows)
    scale, l1, x0 = 1 / (2 * n_rows), 0.03, np.zeros(n_cols)
    lo, hi = zd.shard_bounds(n_rows, rank, world)
    # the exchange of [A^T r | sum r^2]: folded into the kernels over NVLink peer memory
    # (default on one node), and the NCCL all-reduce between the stages (ZF_LASSO_P2P=0)
    sharded = DenseLasso(A[lo:hi], b[lo:hi], l1, scale=scale, distributed=True)
    os.environ["ZF_LASSO_P2P"] = "0"
    sharded_nccl = DenseLasso(A[lo:hi], b[lo:hi], l1, scale=scale, distributed=True)
    del os.environ["ZF_LASSO_P2P"]
    assert not sharded_nccl.peer_exchange
    kinds = [None] * world
    dist.all_gather_object(kinds, sharded.peer_exchange)
    assert len(set(kinds)) == 1, kinds            # every rank took the same decision
    single = DenseLasso(A, b, l1, scale=scale)
    spec = zo.make_least_squares_l1(A, b, l1, scale=scale)
    for opts in (dict(nesterov=True), dict(nesterov=False, max_iter=80),
                 dict(nesterov=True, lr=0.2, decay_rate=1, nesterov_ratio=(0.25, 1 / 64)),
                 dict(nesterov=True, return_all=True), dict(nesterov=True, lr=1e9, max_backtrack_iter=3)):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            r_1 = single.minimize_proximal_gradient(x0, **opts)
            ref = zo.minimize_proximal_gradient(spec, x0, **opts)
        for prob in (sharded, sharded_nccl, sharded):     # (the first one again: repeated solves)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                r_sh = prob.minimize_proximal_gradient(x0, **opts)
            assert r_sh.nit == r_1.nit == ref["nit"], (r_sh.nit, r_1.nit, ref["nit"])
            assert r_sh.status == r_1.status == ref["status"]
            np.testing.assert_allclose(r_sh.x, r_1.x, rtol=1e-9, atol=1e-10)
            np.testing.assert_allclose(r_sh.x, ref["x"], rtol=1e-8, atol=1e-9)
            np.testing.assert_allclose(r_sh.fun, ref["fun"], rtol=1e-9)
            if opts.get("return_all"):
                np.testing.assert_allclose(r_sh.allfuns, np.ravel(ref["allfuns"]), rtol=1e-9)
                np.testing.assert_allclose(r_sh.allerrs, ref["allerrs"], rtol=1e-6, atol=1e-12)
            # all ranks hold the same replicated x
            xs = [None] * world
            dist.all_gather_object(xs, r_sh.x)
            for other in xs[1:]:
                np.testing.assert_array_equal(other, xs[0])
    peer = sharded.peer_exchange
    grid = [(0.0, 0.25), (0.5, 1 / 16), (0.25, 17 / 128), (0.0, 0.0), (0.75, 0.25)]
    X0m = np.random.RandomState(5).standard_normal((len(grid), n_cols)) * 0.1
    msh = DenseLassoMulti(A[lo:hi], b[lo:hi], l1, len(grid), scale=scale, distributed=True)
    for opts in (dict(nesterov=True), dict(nesterov=True, lr=0.2, decay_rate=1),
                 dict(nesterov=False, lr=1e6, max_backtrack_iter=4), dict(nesterov=True, max_iter=7)):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            # device-decided rounds (default), then the host-decided rounds of round 1: bit-equal
            got = msh.minimize_proximal_gradient_batched(X0m, grid, **opts)
            os.environ["ZF_LASSO_HOSTLOOP"] = "1"
            try:
                got_host = msh.minimize_proximal_gradient_batched(X0m, grid, **opts)
            finally:
                del os.environ["ZF_LASSO_HOSTLOOP"]
            for k, ab in enumerate(grid):
                ref = zo.minimize_proximal_gradient(spec, X0m[k], nesterov_ratio=ab, **opts)
                assert got[k].nit == got_host[k].nit == ref["nit"], (k, got[k].nit, got_host[k].nit, ref["nit"])
                assert got[k].status == got_host[k].status == ref["status"]
                np.testing.assert_array_equal(got[k].x, got_host[k].x)
                assert got[k].fun == got_host[k].fun and got[k].lr == got_host[k].lr
                np.testing.assert_allclose(got[k].x, ref["x"], rtol=1e-8, atol=1e-9)
                np.testing.assert_allclose(got[k].fun, ref["fun"], rtol=1e-9)
    prob = zp.JOS1(n_features=20)
    X0 = np.random.RandomState(3).uniform(-2, 4, size=(101, 20))
    full = zd.minimize_proximal_gradient_sharded(prob, X0, nesterov=True, tol_internal=1e-11)
    one = prob.minimize_proximal_gradient_batched(X0, nesterov=True, tol_internal=1e-11)
    np.testing.assert_array_equal(full.nit, one.nit)
    np.testing.assert_array_equal(full.x, one.x)
    dist.barrier()
    if rank == 0:
        print(f"MULTIGPU_OK world={world} peer_exchange={peer}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
