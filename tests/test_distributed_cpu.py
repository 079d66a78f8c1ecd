"""CPU tests (gloo, world_size 2) of the multi-GPU host logic in zfista_b200.distributed:
start sharding with no data-path collective, and the row-sharded LASSO protocol whose one
exchange is the all-reduce of [A^T r | sum r^2].  The local compute is injected (the CPU
oracle / a numpy model of the split protocol), the sharding + exchange code is the
product's."""
import os
import socket
import warnings

import numpy as np
import pytest

import helpers
from zfista_b200 import distributed as zd


def test_shard_bounds_partition_everything_exactly_once():
    for n in (0, 1, 7, 8, 1000, 1024, 15 * 1024):
        for world in (1, 2, 3, 4, 8):
            spans = [zd.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        zd.shard_bounds(4, 2, 2)


def test_momentum_grid_layout():
    X0 = np.arange(6.0).reshape(3, 2)
    Xr, AB, si, gi = zd.momentum_grid(X0, helpers.AB_GRID)
    G = len(helpers.AB_GRID)
    assert Xr.shape == (3 * G, 2) and AB.shape == (3 * G, 2)
    for s in range(3):
        for g in range(G):
            row = s * G + g
            assert si[row] == s and gi[row] == g
            np.testing.assert_array_equal(Xr[row], X0[s])
            np.testing.assert_array_equal(AB[row], helpers.AB_GRID[g])


# ----------------------------------------------------------------------------- workers
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_local_solver(problem, X0, nesterov_ratio=(0, 0.25), **kw):
    """Stand-in for the CUDA batched solve: the CPU oracle, returning a BatchResult."""
    from oracle import zfista_oracle as zo
    from zfista_b200.proximal_gradient import BatchResult

    spec = zo.make_spec(type(problem).__name__, n_features=problem.n_features)
    ab = np.asarray(nesterov_ratio, dtype=np.float64)
    rows = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(len(X0)):
            pair = tuple(ab[i]) if ab.ndim == 2 else tuple(ab)
            rows.append(zo.minimize_proximal_gradient(spec, X0[i], nesterov_ratio=pair, **kw))
    n, m = problem.n_features, problem.n_objectives
    S = len(rows)
    return BatchResult(
        x=np.array([r["x"] for r in rows]).reshape(S, n),
        fun=np.array([r["fun"] for r in rows]).reshape(S, m),
        nit=np.array([r["nit"] for r in rows], dtype=np.int64),
        status=np.array([r["status"] for r in rows], dtype=np.int32),
        lr=np.array([r["lr"] for r in rows], dtype=np.float64),
        nfev=np.zeros(S, dtype=np.int64), n_dual=np.zeros(S, dtype=np.int64),
        err=np.zeros(S), time=0.0)


def _worker_starts(rank, world, port, out_dir):
    import torch.distributed as dist

    import zfista_b200.problems as zp

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank,
                            world_size=world)
    prob = zp.JOS1(n_features=5)
    rng = np.random.RandomState(0)
    X0 = rng.uniform(-2, 4, size=(7, 5))          # 7 starts over 2 ranks: ragged shards
    Xg, AB, _, _ = zd.momentum_grid(X0[:3], helpers.AB_GRID[:3])
    full = zd.minimize_proximal_gradient_sharded(prob, X0, nesterov=True, tol_internal=1e-11,
                                                 local_solver=_oracle_local_solver)
    mine = zd.minimize_proximal_gradient_sharded(prob, X0, nesterov=True, tol_internal=1e-11,
                                                 gather=False, local_solver=_oracle_local_solver)
    grid = zd.minimize_proximal_gradient_sharded(prob, Xg, nesterov_ratio=AB, nesterov=True,
                                                 tol_internal=1e-11,
                                                 local_solver=_oracle_local_solver)
    np.savez(os.path.join(out_dir, f"starts_{rank}.npz"), x=full.x, nit=full.nit, fun=full.fun,
             mine_x=mine.x, grid_nit=grid.nit, grid_x=grid.x)
    dist.destroy_process_group()


def test_sharded_starts_world2_matches_single_process(tmp_path):
    import torch.multiprocessing as mp

    import zfista_b200.problems as zp

    port = _free_port()
    mp.spawn(_worker_starts, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    prob = zp.JOS1(n_features=5)
    rng = np.random.RandomState(0)
    X0 = rng.uniform(-2, 4, size=(7, 5))
    ref = _oracle_local_solver(prob, X0, nesterov=True, tol_internal=1e-11)
    Xg, AB, _, _ = zd.momentum_grid(X0[:3], helpers.AB_GRID[:3])
    ref_grid = _oracle_local_solver(prob, Xg, nesterov_ratio=AB, nesterov=True, tol_internal=1e-11)
    outs = [np.load(tmp_path / f"starts_{r}.npz") for r in range(2)]
    for r, o in enumerate(outs):
        np.testing.assert_array_equal(o["nit"], ref.nit)          # gathered, original order
        np.testing.assert_array_equal(o["x"], ref.x)
        np.testing.assert_array_equal(o["fun"], ref.fun)
        lo, hi = zd.shard_bounds(7, r, 2)
        np.testing.assert_array_equal(o["mine_x"], ref.x[lo:hi])  # gather=False: own slice only
        np.testing.assert_array_equal(o["grid_nit"], ref_grid.nit)
        np.testing.assert_array_equal(o["grid_x"], ref_grid.x)


class NumpySplitLasso:
    """numpy model of the zf_lasso_begin / grad / partial / step / finish state machine
    (csrc/zf_lasso.cu) on ONE row shard; `partial` = [A^T r | sum r^2] is what gets reduced."""

    def __init__(self, A, b, scale, l1, x0, opts):
        self.A, self.b, self.scale, self.l1 = A, b, scale, l1
        self.o = dict(lr=1.0, tol=1e-5, tol_internal=1e-12, max_iter=1000000,
                      max_backtrack_iter=100, decay_rate=0.5, nesterov=False,
                      nesterov_ratio=(0, 0.25), deprecated=False)
        self.o.update(opts)
        self.x0 = x0
        self.partial = np.zeros(A.shape[1] + 1)

    def _f(self, ss):
        return np.sqrt(ss) ** 2 * self.scale

    def begin(self):
        self.xp = self.xn = self.y = self.x0.copy()
        self.lr, self.t, self.nit, self.phase = self.o["lr"], 1.0, 0, "init"
        r = self.A @ self.xp - self.b
        self.partial[-1] = r @ r

    def grad(self, which):
        v = self.y if which == 0 else self.xn
        r = self.A @ v - self.b
        if which == 0:
            self.partial[:-1] = self.A.T @ r
        self.partial[-1] = r @ r

    def _trial(self):
        self.xn = np.sign(self.y - self.lr * self.g) * np.maximum(
            np.abs(self.y - self.lr * self.g) - self.lr * self.l1, 0)
        d = self.xn - self.y
        self.gx = self.l1 * np.abs(self.xn).sum()
        self.sub = self.g @ d + self.gx + np.sqrt(d @ d) ** 2 / 2 / self.lr
        if not self.o["deprecated"]:
            self.sub += self.f_y - self.F_prev
        self.err = np.max(np.abs(d))

    def _accept(self):
        if self.err < self.o["tol"] or self.nit >= self.o["max_iter"]:
            self.status = 1 if self.err < self.o["tol"] else 0
            self.phase = "done"
            return 2
        mom = 0.0
        if self.o["nesterov"]:
            a, b = self.o["nesterov_ratio"]
            t_new = np.sqrt(self.t ** 2 - a * self.t + b) + 0.5
            mom, self.t = (self.t - 1) / t_new, t_new
        self.y = self.xn + mom * (self.xn - self.xp)
        self.xp, self.F_prev, self.nit, self.phase = self.xn, self.F_x, self.nit + 1, "grad"
        return 0

    def step(self):
        ss = self.partial[-1]
        if self.phase == "init":
            self.F_prev = self.F_x = self._f(ss) + self.l1 * np.abs(self.xp).sum()
            self.nit, self.phase = 1, "grad"
            return 0
        if self.phase == "grad":
            self.g = self.partial[:-1] * (2 * self.scale)
            self.f_y = self._f(ss)
            self.bt = 0
            self._trial()
            self.phase = "fnew"
            return 1
        if self.phase == "fnew":
            f_x = self._f(ss)
            self.F_x = f_x + self.gx
            ok = (f_x - self.f_y if self.o["deprecated"] else self.F_x - self.F_prev) \
                <= self.sub + self.o["tol_internal"]
            if ok or self.o["decay_rate"] == 1:
                return self._accept()
            self.lr *= self.o["decay_rate"]
            self.bt += 1
            if self.bt >= self.o["max_backtrack_iter"]:
                self.status, self.phase, self.xn, self.nit = -1, "done", self.xp, self.nit - 1
                return 2
            self._trial()
            return 1
        return 2

    def finish(self):
        return dict(x=self.xn, fun=self.F_x, nit=self.nit, status=self.status)


def _lasso_problem():
    rng = np.random.RandomState(5)
    A = rng.standard_normal((90, 40))
    w = np.zeros(40)
    w[:6] = rng.standard_normal(6)
    b = A @ w + 0.01 * rng.standard_normal(90)
    return A, b, 1 / 180, 0.05, np.zeros(40)


def _worker_lasso(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank,
                            world_size=world)
    A, b, scale, l1, x0 = _lasso_problem()
    lo, hi = zd.shard_bounds(len(A), rank, world)
    for tag, opts in (("fista", dict(nesterov=True)), ("ista", dict(nesterov=False, max_iter=50))):
        ops = NumpySplitLasso(A[lo:hi], b[lo:hi], scale, l1, x0, opts)
        buf = torch.from_numpy(ops.partial)             # shares memory with ops.partial

        def allreduce():
            dist.all_reduce(buf)

        res = zd.run_split_lasso(ops, allreduce)
        np.savez(os.path.join(out_dir, f"lasso_{tag}_{rank}.npz"), **res)
    dist.destroy_process_group()


def test_row_sharded_lasso_world2_matches_oracle(tmp_path):
    import torch.multiprocessing as mp

    from oracle import zfista_oracle as zo

    port = _free_port()
    mp.spawn(_worker_lasso, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    A, b, scale, l1, x0 = _lasso_problem()
    spec = zo.make_least_squares_l1(A, b, l1, scale=scale)
    for tag, opts in (("fista", dict(nesterov=True)), ("ista", dict(nesterov=False, max_iter=50))):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = zo.minimize_proximal_gradient(spec, x0, **opts)
        outs = [np.load(tmp_path / f"lasso_{tag}_{r}.npz") for r in range(2)]
        np.testing.assert_array_equal(outs[0]["x"], outs[1]["x"])      # ranks agree bit for bit
        for o in outs:
            assert int(o["nit"]) == ref["nit"]
            assert int(o["status"]) == ref["status"]
            np.testing.assert_allclose(o["x"], ref["x"], rtol=1e-9, atol=1e-10)
            np.testing.assert_allclose(float(o["fun"]), ref["fun"], rtol=1e-10)
        # one process, one shard == the same protocol without an exchange
        solo = zd.run_split_lasso(NumpySplitLasso(A, b, scale, l1, x0, opts), lambda: None)
        assert solo["nit"] == ref["nit"]
        np.testing.assert_allclose(solo["x"], ref["x"], rtol=1e-9, atol=1e-10)


class NumpyDeviceLasso:
    """numpy model of the DEVICE-decided protocol (zf_lasso_dev_* in csrc/zf_lasso.cu) on ONE row
    shard: stages read and write a state dict exactly as the kernels read and write the device
    struct, stages enqueued after the end do nothing, and the host only sees snapshots."""

    def __init__(self, A, b, scale, l1, x0, opts):
        self.A, self.b, self.scale, self.l1, self.x0 = A, b, scale, l1, x0
        self.o = dict(lr=1.0, tol=1e-5, tol_internal=1e-12, max_iter=1000000,
                      max_backtrack_iter=100, decay_rate=0.5, nesterov=False,
                      nesterov_ratio=(0, 0.25), deprecated=False)
        self.o.update(opts)
        self.partial = np.zeros(A.shape[1] + 1)
        self.snaps = [None, None]
        self.stages_after_done = 0

    def _f(self, ss):
        return np.sqrt(ss) ** 2 * self.scale

    def needs_feval(self):
        return self.o["decay_rate"] != 1

    def begin(self):
        self.xp, self.xn, self.y = self.x0.copy(), self.x0.copy(), self.x0.copy()
        r = self.A @ self.xp - self.b
        self.partial[-1] = r @ r
        self.s = dict(done=False)

    def _next_momentum(self):
        if not self.o["nesterov"]:
            return self.s["t"], 0.0
        a, b = self.o["nesterov_ratio"]
        t = self.s["t"]
        t_new = np.sqrt(t * t - a * t + b) + 0.5
        return t_new, (t - 1) / t_new

    def _accept(self, maxd):
        s = self.s
        s["err"] = maxd
        if maxd < self.o["tol"] or s["nit"] >= self.o["max_iter"]:
            s.update(status=1 if maxd < self.o["tol"] else 0, done=True, skip_grad=True,
                     accept=False, result_is_prev=False)
            return False
        s["t"], s["mom"] = self._next_momentum()
        s.update(F_prev=s["F_x"], nit=s["nit"] + 1, phase="grad", skip_grad=False, bt=0)
        return True

    def stage(self, k):
        s = self.s
        if s["done"] and k in (1, 2, 3, 4):
            self.stages_after_done += 1
        if k == 0:
            F0 = self._f(self.partial[-1]) + self.l1 * np.abs(self.xp).sum()
            s.update(lr=self.o["lr"], t=1.0, F_prev=F0, F_x=F0, nit=1, status=0, phase="grad",
                     bt=0, accept=False, skip_grad=False, F_known=False, result_is_prev=False)
        elif k == 1:
            if not s["skip_grad"]:
                r = self.A @ self.y - self.b
                self.gpart, self.sq = self.A.T @ r, r @ r
            self.partial[:-1], self.partial[-1] = self.gpart, self.sq     # collect
        elif k == 2:
            if s["done"]:
                return
            retry = s["phase"] == "retry"
            if not retry:
                self.g = self.partial[:-1] * (2 * self.scale)
            lr = s["lr"]
            v = self.y - lr * self.g
            xn = np.sign(v) * np.maximum(np.abs(v) - lr * self.l1, 0)
            d = xn - self.y
            self.xn = xn
            s.update(abs1=np.abs(xn).sum(), maxd=np.max(np.abs(d)), accept=False)
            if not self.needs_feval():
                _, mom = self._next_momentum()
                self.y = xn + mom * (xn - self.xp)
                self.xp = xn.copy()
                s["F_known"] = False
                self._accept(s["maxd"])
            else:
                if not retry:
                    s["f_y"] = self._f(self.partial[-1])
                sub = self.g @ d + self.l1 * s["abs1"] + np.sqrt(d @ d) ** 2 / 2 / lr
                if not self.o["deprecated"]:
                    sub += s["f_y"] - s["F_prev"]
                s["sub"] = sub
        elif k in (3, 6):
            if k == 3 and s["done"]:
                return
            r = self.A @ self.xn - self.b
            self.partial[-1] = r @ r
        elif k == 4:
            if s["done"]:
                return
            f_x = self._f(self.partial[-1])
            s["F_x"], s["F_known"] = f_x + self.l1 * s["abs1"], True
            ok = (f_x - s["f_y"] if self.o["deprecated"] else s["F_x"] - s["F_prev"]) \
                <= s["sub"] + self.o["tol_internal"]
            if ok:
                if self._accept(s["maxd"]):
                    self.y = self.xn + s["mom"] * (self.xn - self.xp)
                    self.xp = self.xn.copy()
                return
            s["lr"] *= self.o["decay_rate"]
            s["bt"] += 1
            if s["bt"] >= self.o["max_backtrack_iter"]:
                s.update(result_is_prev=True, F_x=s["F_prev"], nit=s["nit"] - 1, status=-1,
                         done=True, skip_grad=True)
            else:
                s.update(phase="retry", skip_grad=True)
        elif k == 5:
            if not s["F_known"] and not s["result_is_prev"]:
                s["F_x"] = self._f(self.partial[-1]) + self.l1 * s["abs1"]

    def snapshot(self, slot):
        self.snaps[slot] = self.s["done"]

    def wait(self, slot):
        return self.snaps[slot]

    def finish(self):
        s = self.s
        assert s["done"]
        return dict(x=self.xp if s["result_is_prev"] else self.xn, fun=s["F_x"], nit=s["nit"],
                    status=s["status"], wasted=self.stages_after_done)


DEVICE_CASES = (("fista", dict(nesterov=True)), ("ista", dict(nesterov=False, max_iter=50)),
                ("fixed", dict(nesterov=True, lr=0.5, decay_rate=1, max_iter=200)),
                ("fail", dict(nesterov=True, lr=1e6, max_backtrack_iter=3)))


def _worker_device_lasso(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank,
                            world_size=world)
    A, b, scale, l1, x0 = _lasso_problem()
    lo, hi = zd.shard_bounds(len(A), rank, world)
    for tag, opts in DEVICE_CASES:
        ops = NumpyDeviceLasso(A[lo:hi], b[lo:hi], scale, l1, x0, opts)
        buf = torch.from_numpy(ops.partial)             # shares memory with ops.partial
        res = zd.run_device_lasso(ops, lambda: dist.all_reduce(buf),
                                  lambda: dist.all_reduce(buf[-1:]), chunk=5)
        np.savez(os.path.join(out_dir, f"dlasso_{tag}_{rank}.npz"), **res)
    dist.destroy_process_group()


def test_device_decided_lasso_world2_matches_oracle(tmp_path):
    """The device-decided protocol (no host decision inside the loop, the host polling one chunk
    behind) over two row shards: the same nit / status / x / F as the oracle and as the
    host-decided protocol, including a line-search failure and a fixed-step run."""
    import torch.multiprocessing as mp

    from oracle import zfista_oracle as zo

    port = _free_port()
    mp.spawn(_worker_device_lasso, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    A, b, scale, l1, x0 = _lasso_problem()
    spec = zo.make_least_squares_l1(A, b, l1, scale=scale)
    for tag, opts in DEVICE_CASES:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = zo.minimize_proximal_gradient(spec, x0, **opts)
        outs = [np.load(tmp_path / f"dlasso_{tag}_{r}.npz") for r in range(2)]
        np.testing.assert_array_equal(outs[0]["x"], outs[1]["x"])      # ranks agree bit for bit
        for o in outs:
            assert int(o["nit"]) == ref["nit"], (tag, int(o["nit"]), ref["nit"])
            assert int(o["status"]) == ref["status"]
            np.testing.assert_allclose(o["x"], ref["x"], rtol=1e-9, atol=1e-10)
            np.testing.assert_allclose(float(o["fun"]), ref["fun"], rtol=1e-10)
            # at most two chunks of 5 trials (x up to 4 stages) run past the end
            assert int(o["wasted"]) <= 2 * 5 * 4
        solo = zd.run_device_lasso(NumpyDeviceLasso(A, b, scale, l1, x0, opts), lambda: None,
                                   lambda: None, chunk=3)
        assert solo["nit"] == ref["nit"] and solo["status"] == ref["status"]
        np.testing.assert_allclose(solo["x"], ref["x"], rtol=1e-9, atol=1e-10)


class NumpyDeviceLassoMulti:
    """numpy model of the device-decided LOCKSTEP rounds of K runs sharing A (zf_lasso_multi_dev_*
    in csrc/zf_lasso_multi.cu) on ONE row shard: per-run scalars and the run masks (active / on
    trial / accepted / advancing) as the one-warp kernels keep them, a gradient stage that does
    nothing while a retry is pending (so that the unconditional all-reduce after it re-reduces
    stale values nobody reads), stages after the end that do nothing."""

    def __init__(self, A, b, scale, l1, X0, ab, opts):
        self.A, self.b, self.scale, self.l1 = A, b, scale, l1
        self.X0, self.ab = np.array(X0, dtype=float), np.asarray(ab, dtype=float)
        self.K, self.n = self.X0.shape
        self.o = dict(lr=1.0, tol=1e-5, tol_internal=1e-12, max_iter=1000000,
                      max_backtrack_iter=100, decay_rate=0.5, nesterov=False, deprecated=False)
        self.o.update(opts)
        self.partial = np.zeros(self.K * self.n + self.K)      # [K][n] gradients | K residual norms
        self.snaps = [None, None]

    # views into `partial`
    def _g(self):
        return self.partial[:self.K * self.n].reshape(self.K, self.n)

    def _ss(self):
        return self.partial[self.K * self.n:]

    def _f(self, ss):
        return np.sqrt(ss) ** 2 * self.scale

    def needs_feval(self):
        return self.o["decay_rate"] != 1

    def _residual_norms(self, X):
        R = X @ self.A.T - self.b
        self._ss()[:] = np.einsum("kr,kr->k", R, R)
        return R

    def begin(self):
        self.Xp, self.Xn, self.Y = self.X0.copy(), self.X0.copy(), self.X0.copy()
        self.G = np.zeros_like(self.X0)
        self._residual_norms(self.Y)
        self.s = dict(done=False)

    def _advance(self, accepted):
        s, o = self.s, self.o
        stop, adv = set(), set()
        for k in accepted:
            r = s["run"][k]
            r["err"] = r["maxd"]
            if r["err"] < o["tol"] or r["nit"] >= o["max_iter"]:
                r["status"] = 1 if r["err"] < o["tol"] else 0
                stop.add(k)
            else:
                mom = 0.0
                if o["nesterov"]:
                    a, b = self.ab[k]
                    t = r["t"]
                    tn = np.sqrt(t * t - a * t + b) + 0.5
                    mom, r["t"] = (t - 1) / tn, tn
                r.update(mom=mom, F_prev=r["F_x"], nit=r["nit"] + 1)
                adv.add(k)
        s["active"] -= stop
        s.update(adv=adv, accepted=set(), trial=set())
        if s["active"]:
            s.update(phase="grad", skip_grad=False)
        else:
            s.update(phase="done", skip_grad=True, done=True)

    def _momentum(self):
        for k in self.s["adv"]:
            xn = self.Xn[k]
            self.Y[k] = xn + self.s["run"][k]["mom"] * (xn - self.Xp[k])
            self.Xp[k] = xn.copy()

    def stage(self, st):
        s, o = self.s, self.o
        if st == 0:
            ss = self._ss()
            s["run"] = []
            for k in range(self.K):
                abs1 = np.abs(self.X0[k]).sum()
                F0 = self._f(ss[k]) + self.l1 * abs1
                s["run"].append(dict(lr=o["lr"], t=1.0, F_prev=F0, F_x=F0, f_y=0.0, sub=0.0,
                                     err=np.inf, mom=0.0, abs1=abs1, maxd=0.0, gd=0.0, dd=0.0,
                                     nit=1, status=0, bt=0, F_known=False, result_is_prev=False))
            s.update(active=set(range(self.K)), trial=set(), accepted=set(), adv=set(),
                     phase="grad", skip_grad=False, done=False)
        elif st == 1:
            if s["skip_grad"]:
                return                                   # `partial` keeps whatever it holds
            R = self._residual_norms(self.Y)
            self._g()[:] = R @ self.A
        elif st == 2:
            if s["done"]:
                return                                   # (adv is empty once every run has stopped)
            first = s["phase"] == "grad"
            mask = set(s["active"] if first else s["trial"])
            for k in mask:
                r = s["run"][k]
                if first:
                    self.G[k] = self._g()[k] * (2 * self.scale)
                g, lr, y = self.G[k], r["lr"], self.Y[k]
                v = y - lr * g
                xn = np.sign(v) * np.maximum(np.abs(v) - lr * self.l1, 0)
                d = xn - y
                self.Xn[k] = xn
                r.update(gd=g @ d, dd=d @ d, abs1=np.abs(xn).sum(), maxd=np.max(np.abs(d)))
                if first:
                    r.update(bt=0, f_y=self._f(self._ss()[k]), F_known=False)
                sub = r["gd"] + self.l1 * r["abs1"] + np.sqrt(r["dd"]) ** 2 / 2 / lr
                if not o["deprecated"]:
                    sub += r["f_y"] - r["F_prev"]
                r["sub"] = sub
            if not self.needs_feval():
                self._advance(mask)
                self._momentum()
            else:
                s.update(trial=mask, adv=set())
        elif st in (3, 6):
            if st == 3 and s["done"]:
                return
            self._residual_norms(self.Xn)
        elif st == 4:
            if not s["done"]:
                acc, fail = set(), set()
                for k in s["trial"]:
                    r = s["run"][k]
                    f_x = self._f(self._ss()[k])
                    r["F_x"], r["F_known"] = f_x + self.l1 * r["abs1"], True
                    if o["decay_rate"] == 1:
                        ok = True
                    elif o["deprecated"]:
                        ok = f_x - r["f_y"] <= r["sub"] + o["tol_internal"]
                    else:
                        ok = r["F_x"] - r["F_prev"] <= r["sub"] + o["tol_internal"]
                    if ok:
                        acc.add(k)
                        continue
                    r["lr"] *= o["decay_rate"]
                    r["bt"] += 1
                    if r["bt"] >= o["max_backtrack_iter"]:
                        r.update(result_is_prev=True, F_x=r["F_prev"], nit=r["nit"] - 1, status=-1)
                        fail.add(k)
                left = s["trial"] - acc - fail
                accepted = s["accepted"] | acc
                s["active"] -= fail
                s.update(accepted=accepted, trial=left)
                if left:
                    s.update(phase="retry", skip_grad=True, adv=set())
                else:
                    self._advance(accepted)
            self._momentum()
        elif st == 5:
            for k, r in enumerate(s["run"]):
                if not r["F_known"] and not r["result_is_prev"]:
                    r["F_x"] = self._f(self._ss()[k]) + self.l1 * r["abs1"]

    def snapshot(self, slot):
        self.snaps[slot] = self.s["done"]

    def wait(self, slot):
        return self.snaps[slot]

    def finish(self):
        s = self.s
        assert s["done"]
        runs = s["run"]
        x = np.stack([self.Xp[k] if r["result_is_prev"] else self.Xn[k] for k, r in enumerate(runs)])
        return dict(x=x, fun=np.array([r["F_x"] for r in runs]),
                    nit=np.array([r["nit"] for r in runs]),
                    status=np.array([r["status"] for r in runs]),
                    lr=np.array([r["lr"] for r in runs]))


MULTI_AB = [(0.0, 0.25), (0.5, 1 / 16), (0.25, 17 / 128), (0.0, 0.0), (0.75, 0.25)]
MULTI_CASES = (("fista", dict(nesterov=True)), ("ista", dict(nesterov=False, max_iter=40)),
               ("fixed", dict(nesterov=True, lr=0.5, decay_rate=1, max_iter=300)),
               ("fail", dict(nesterov=True, lr=1e6, max_backtrack_iter=3)),
               ("short", dict(nesterov=True, max_iter=7)))


def _multi_starts():
    return np.random.RandomState(8).standard_normal((len(MULTI_AB), 40)) * 0.2


def _worker_device_lasso_multi(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank,
                            world_size=world)
    A, b, scale, l1, _ = _lasso_problem()
    lo, hi = zd.shard_bounds(len(A), rank, world)
    X0 = _multi_starts()
    for tag, opts in MULTI_CASES:
        ops = NumpyDeviceLassoMulti(A[lo:hi], b[lo:hi], scale, l1, X0, MULTI_AB, opts)
        buf = torch.from_numpy(ops.partial)             # shares memory with ops.partial
        res = zd.run_device_lasso(ops, lambda: dist.all_reduce(buf),
                                  lambda: dist.all_reduce(buf[ops.K * ops.n:]), chunk=4)
        np.savez(os.path.join(out_dir, f"mlasso_{tag}_{rank}.npz"), **res)
    dist.destroy_process_group()


def test_device_decided_multi_run_lasso_world2_matches_oracle(tmp_path):
    """The lockstep rounds of several runs sharing A, decided "on the device", over two row
    shards with the exchange after the gradient and the F stages: run by run the oracle's nit /
    status / x / F (runs retry their line searches and stop at different rounds; one case fails
    its line search, one stops at max_iter), ranks bit-equal, and the same through one shard."""
    import torch.multiprocessing as mp

    from oracle import zfista_oracle as zo

    port = _free_port()
    mp.spawn(_worker_device_lasso_multi, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    A, b, scale, l1, _ = _lasso_problem()
    X0 = _multi_starts()
    spec = zo.make_least_squares_l1(A, b, l1, scale=scale)
    for tag, opts in MULTI_CASES:
        outs = [np.load(tmp_path / f"mlasso_{tag}_{r}.npz") for r in range(2)]
        np.testing.assert_array_equal(outs[0]["x"], outs[1]["x"])
        solo = zd.run_device_lasso(NumpyDeviceLassoMulti(A, b, scale, l1, X0, MULTI_AB, opts),
                                   lambda: None, lambda: None, chunk=3)
        for k, ab in enumerate(MULTI_AB):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ref = zo.minimize_proximal_gradient(spec, X0[k], nesterov_ratio=ab, **opts)
            for o in (outs[0], outs[1], solo):
                assert int(o["nit"][k]) == ref["nit"], (tag, k, int(o["nit"][k]), ref["nit"])
                assert int(o["status"][k]) == ref["status"], (tag, k)
                np.testing.assert_allclose(o["x"][k], ref["x"], rtol=1e-9, atol=1e-10)
                np.testing.assert_allclose(float(o["fun"][k]), ref["fun"], rtol=1e-10)
