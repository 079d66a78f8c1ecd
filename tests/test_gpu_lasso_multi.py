"""GPU parity tests of the shared-A multi-run LASSO path (DESIGN.md row c''):
csrc/zf_lasso_multi.cu (FP64 tensor-core DGEMM passes) through
zfista_b200.lasso.DenseLassoMulti, against numpy, the CPU oracle and the single-run path."""
import warnings

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu
REL = 1e-8


def _close(a, b, rel=REL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    np.testing.assert_allclose(a, b, rtol=rel, atol=rel * max(1.0, float(np.max(np.abs(b)))))


@pytest.mark.parametrize("n_rows,n_cols,n_runs,batched_b", [
    (1, 2, 1, False), (3, 2, 2, True), (7, 6, 5, False), (50, 200, 8, True),
    (257, 1022, 13, False), (1000, 2050, 16, True), (641, 130, 24, True),
    (3001, 4098, 32, False), (129, 64, 9, True), (64, 128, 17, False)])
def test_gradients_match_numpy(gpu, n_rows, n_cols, n_runs, batched_b):
    """grad_k = 2*scale*A^T(A x_k - b_k), f_k = scale*||A x_k - b_k||^2 for every run: ragged
    row / column counts (partial tiles in both passes), every n-tile count, shared and
    per-run b."""
    from zfista_b200.lasso import DenseLassoMulti

    rng = np.random.RandomState(n_rows + n_cols + n_runs)
    A = rng.standard_normal((n_rows, n_cols))
    b = rng.standard_normal((n_runs, n_rows)) if batched_b else rng.standard_normal(n_rows)
    X = rng.standard_normal((n_runs, n_cols))
    prob = DenseLassoMulti(A, b, 0.1, n_runs, scale=0.37)
    for _ in range(2):                       # second call: ring barriers restart cleanly
        grad, f = prob.gradient(X)
        Rm = X @ A.T - b                     # (n_runs, n_rows)
        _close(grad.cpu().numpy(), Rm @ A * (2 * 0.37), rel=1e-12)
        _close(f.cpu().numpy(), np.sum(Rm * Rm, axis=1) * 0.37, rel=1e-12)
    # one x shared by all runs
    grad, f = prob.gradient(X[0])
    Rm = (A @ X[0])[None, :] - b
    _close(grad.cpu().numpy(), np.broadcast_to(Rm, (n_runs, n_rows)) @ A * (2 * 0.37), rel=1e-12)


def test_odd_columns_are_refused(gpu):
    from zfista_b200 import _lib
    from zfista_b200.lasso import DenseLassoMulti

    with pytest.raises(_lib.ZfError) as e:
        DenseLassoMulti(np.ones((4, 3)), np.ones(4), 0.1, 2)
    assert e.value.code == -3
    with pytest.raises(ValueError):
        DenseLassoMulti(np.ones((4, 4)), np.ones(4), 0.1, 33)
    with pytest.raises(ValueError):
        DenseLassoMulti(np.ones((4, 4)), np.ones((3, 4)), 0.1, 2)


def _dataset(seed, n_rows, n_cols, n_runs, batched_b):
    rng = np.random.RandomState(seed)
    A = rng.standard_normal((n_rows, n_cols))
    w = np.zeros((n_runs if batched_b else 1, n_cols))
    w[:, :10] = rng.standard_normal((len(w), 10))
    b = w @ A.T + 0.01 * rng.standard_normal((len(w), n_rows))
    X0 = rng.standard_normal((n_runs, n_cols)) * 0.1
    return A, (b if batched_b else b[0]), X0


@pytest.mark.parametrize("n_rows,n_cols,batched_b", [(300, 120, False), (200, 502, True)])
def test_runs_match_oracle_run_by_run(gpu, n_rows, n_cols, batched_b):
    """Every run of a lockstep call == the reference algorithm run alone on that run's
    (x0, b, (a, b)): same nit, x and F within 1e-8 (north_star's tolerance), with
    backtracking (runs retry and finish at different rounds), deprecated, ISTA, fixed step."""
    from oracle import zfista_oracle as zo
    from zfista_b200.lasso import DenseLassoMulti

    n_runs = len(helpers.AB_GRID)
    A, b, X0 = _dataset(n_rows + n_cols, n_rows, n_cols, n_runs, batched_b)
    scale, l1 = 1 / (2 * n_rows), 0.05
    prob = DenseLassoMulti(A, b, l1, n_runs, scale=scale)
    L = 2 * scale * np.linalg.norm(A, 2) ** 2
    for opts in (dict(nesterov=True), dict(nesterov=False, max_iter=120),
                 dict(nesterov=True, deprecated=True),
                 dict(nesterov=True, lr=1 / L, decay_rate=1), dict(nesterov=True, lr=4.0)):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            res = prob.minimize_proximal_gradient_batched(X0, helpers.AB_GRID, return_all=True,
                                                          **opts)
            nits = set()
            for k, ab in enumerate(helpers.AB_GRID):
                spec = zo.make_least_squares_l1(A, b[k] if batched_b else b, l1, scale=scale)
                ref = zo.minimize_proximal_gradient(spec, X0[k], nesterov_ratio=ab,
                                                    return_all=True, **opts)
                assert res[k].nit == ref["nit"] and res[k].success == ref["success"], (opts, k)
                _close(res[k].x, ref["x"])
                _close(res[k].fun, ref["fun"])
                _close(res[k].allerrs, ref["allerrs"], rel=1e-6)
                _close(np.ravel(res[k].allfuns), np.ravel(ref["allfuns"]))
                nits.add(res[k].nit)
        if opts.get("nesterov") and "max_iter" not in opts:
            assert len(nits) > 1             # the runs really stop at different rounds
        # without traces a fixed-step call skips the F passes but lands on the same iterates
        if opts.get("decay_rate") == 1:
            res2 = prob.minimize_proximal_gradient_batched(X0, helpers.AB_GRID, **opts)
            for r2, r1 in zip(res2, res):
                assert r2.nit == r1.nit and r2.allerrs is None
                np.testing.assert_array_equal(r2.x, r1.x)
                _close(r2.fun, r1.fun, rel=1e-13)


def test_runs_match_single_run_path(gpu):
    """shared x0 and b, per-run (a, b): each run == DenseLasso (csrc/zf_lasso.cu) solving it
    alone, and == the golden fixture the unmodified reference produced for (0, 1/4)."""
    from zfista_b200.lasso import DenseLasso, DenseLassoMulti

    d = helpers.load("lasso_single")
    A, b = d["ds_A"], d["ds_b"]
    if A.shape[1] % 2:
        A = A[:, :-1]
    l1, scale = float(d["ds_l1"]), float(d["ds_scale"])
    x0 = np.asarray(d["ds_x0"])[:A.shape[1]]
    grid = helpers.AB_GRID[:6]
    multi = DenseLassoMulti(A, b, l1, len(grid), scale=scale)
    single = DenseLasso(A, b, l1, scale=scale)
    res = multi.minimize_proximal_gradient_batched(x0, grid, nesterov=True, max_iter=20000)
    for k, ab in enumerate(grid):
        ref = single.minimize_proximal_gradient(x0, nesterov=True, nesterov_ratio=ab, max_iter=20000)
        assert res[k].nit == ref.nit and res[k].success == ref.success
        _close(res[k].x, ref.x)
        _close(res[k].fun, ref.fun)
    if A.shape[1] == d["ds_A"].shape[1]:
        k = grid.index((0.0, 1 / 4))
        assert res[k].nit == int(d["ds_bt_fista_nit"])
        _close(res[k].x, d["ds_bt_fista_x"])
        _close(res[k].fun, d["ds_bt_fista_fun"])


def test_ab_sweep_matches_the_reference_fixture(gpu):
    """tests/golden/lasso_ab_sweep.npz holds what the UNMODIFIED reference returns for the 15
    (a, b) pairs of the momentum grid on one shared A (one solve per pair, per-run x0), with the
    line search and with a fixed step.  One lockstep call must give every run's nit exactly and
    x, F within 1e-8."""
    from zfista_b200.lasso import DenseLassoMulti

    d = helpers.load("lasso_ab_sweep")
    grid = [(float(a), float(b)) for a, b in d["grid"]]
    prob = DenseLassoMulti(d["A"], d["b"], float(d["l1"]), len(grid), scale=float(d["scale"]))
    L = float(d["L"])
    for tag, opts in {"bt": dict(), "fixed": dict(lr=1 / L, decay_rate=1)}.items():
        res = prob.minimize_proximal_gradient_batched(d["X0"], grid, nesterov=True,
                                                      max_iter=20000, **opts)
        assert [r.nit for r in res] == [int(v) for v in d[f"{tag}_nit"]], tag
        assert all(r.success for r in res)
        for k, r in enumerate(res):
            _close(r.x, d[f"{tag}_x"][k])
            _close(r.fun, d[f"{tag}_fun"][k])


def test_failure_and_max_iter_are_per_run(gpu):
    """Status handling is per run: with a huge initial step and only 2 trials every run fails
    its line search (x = x0, nit = 0, status -1) exactly as a solo solve does; max_iter is hit
    per run; a run that starts at its fixed point stops after one iteration while the others
    go on."""
    from zfista_b200.lasso import DenseLassoMulti

    rng = np.random.RandomState(1)
    A = rng.standard_normal((40, 30)) * 100
    b = rng.standard_normal(40)
    X0 = np.ones((3, 30))
    X0[1] = 0.0
    prob = DenseLassoMulti(A, b, 0.1, 3)
    res = prob.minimize_proximal_gradient_batched(X0, lr=1e6, max_backtrack_iter=2)
    for k in range(3):
        assert res[k].status == -1 and not res[k].success and res[k].nit == 0
        np.testing.assert_array_equal(res[k].x, X0[k])
    res = prob.minimize_proximal_gradient_batched(X0, max_iter=4)
    assert all(r.status == 0 and r.nit == 4 for r in res)
    # a run that starts at its fixed point stops at nit = 1 while the others go on
    prob0 = DenseLassoMulti(A, np.zeros((3, 40)), 0.1, 3)
    res = prob0.minimize_proximal_gradient_batched(X0, max_iter=50)
    assert res[1].nit == 1 and res[1].success and np.all(res[1].x == 0.0)
    assert res[0].nit > 1


def test_large_a_properties(gpu):
    """A (1.07 GB) >> L2: size-independent properties, no CPU pass over A.
    * columns are independent: run k of a 16-run call == run k of a 3-run call;
    * affine in x:  J(x1 + x2) + J(0) == J(x1) + J(x2);
    * the multi gradient equals the single-run kernels' gradient of the same x."""
    import torch

    from zfista_b200.lasso import DenseLasso, DenseLassoMulti

    n_rows, n_cols = 16384 + 37, 8192 + 2
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(n_rows, n_cols, dtype=torch.float64, device="cuda", generator=g) / 64
    b = torch.randn(n_rows, dtype=torch.float64, device="cuda", generator=g)
    X = torch.randn(16, n_cols, dtype=torch.float64, device="cuda", generator=g)
    X[2] = 0.0
    X[3] = X[0] + X[1]
    m16 = DenseLassoMulti(A, b, 0.1, 16, scale=0.5)
    G16, f16 = m16.gradient(X)
    G16, f16 = G16.cpu().numpy(), f16.cpu().numpy()
    m3 = DenseLassoMulti(A, b, 0.1, 3, scale=0.5)
    G3, f3 = m3.gradient(X[:3].contiguous())
    _close(G3.cpu().numpy(), G16[:3], rel=1e-12)
    _close(f3.cpu().numpy(), f16[:3], rel=1e-12)
    _close(G16[3] + G16[2], G16[0] + G16[1], rel=1e-11)
    single = DenseLasso(A, b, 0.1, scale=0.5)
    for k in (0, 5, 15):
        gs, fs = single.gradient(X[k])
        _close(G16[k], gs.cpu().numpy(), rel=1e-11)
        _close(f16[k], fs.item(), rel=1e-12)


def test_device_decided_rounds_equal_host_decided_rounds(gpu):
    """zf_lasso_multi_solve runs the device-decided rounds (per-run scalars and the run masks in
    device memory, one-warp decision kernels, the host polling one chunk of trials behind); the
    split begin / grad / step / finish entry points still decide on the host.  Same bits: nit,
    status, x, F, lr of every run -- runs retrying their line search and stopping at different
    rounds, a fixed step, failures, max_iter, traces."""
    import ctypes as C

    import torch

    from zfista_b200 import _lib
    from zfista_b200.lasso import DenseLassoMulti
    from zfista_b200.proximal_gradient import _make_options

    n_rows, n_cols = 310, 144
    grid = helpers.AB_GRID[:11]
    K = len(grid)
    A, b, X0 = _dataset(77, n_rows, n_cols, K, True)
    scale, l1 = 1 / (2 * n_rows), 0.05
    prob = DenseLassoMulti(A, b, l1, K, scale=scale)
    L = _lib.lib()
    lip = 2 * scale * np.linalg.norm(A, 2) ** 2
    ab = np.ascontiguousarray(np.array(grid, dtype=np.float64))

    def host_rounds(opts):
        o = _make_options(opts.get("lr", 1), opts.get("tol", 1e-5), 1e-12,
                          opts.get("max_iter", 1000000), 100000,
                          opts.get("max_backtrack_iter", 100), False, opts.get("decay_rate", 0.5),
                          opts.get("nesterov", False), (0, 0.25), opts.get("deprecated", False),
                          "reference", 0)
        x0d = torch.from_numpy(X0).cuda()
        xd = torch.empty_like(x0d)
        fun, lrs, err = np.empty(K), np.empty(K), np.empty(K)
        nit, status = np.zeros(K, dtype=np.int64), np.zeros(K, dtype=np.int32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        nxt = C.c_int32(0)
        _lib.check(L.zf_lasso_multi_begin(prob._h, C.byref(o), C.c_void_p(x0d.data_ptr()), 1, p(ab)))
        _lib.check(L.zf_lasso_multi_step(prob._h, C.byref(nxt)))
        while nxt.value != 2:
            _lib.check(L.zf_lasso_multi_grad(prob._h, nxt.value))
            _lib.check(L.zf_lasso_multi_step(prob._h, C.byref(nxt)))
        _lib.check(L.zf_lasso_multi_finish(prob._h, C.c_void_p(xd.data_ptr()), p(fun), p(nit),
                                           p(status), p(lrs), p(err)))
        return xd.cpu().numpy(), fun, nit, status, lrs

    cases = [dict(nesterov=True), dict(nesterov=False, max_iter=30),
             dict(nesterov=True, lr=1 / lip, decay_rate=1, max_iter=250),
             dict(nesterov=True, lr=1 / lip, decay_rate=1, max_iter=19, tol=0.0),
             dict(nesterov=True, deprecated=True, lr=8.0),
             dict(nesterov=True, lr=1e9, max_backtrack_iter=3)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for opts in cases + cases[:2]:
            got = prob.minimize_proximal_gradient_batched(X0, grid, **opts)
            x, fun, nit, status, lrs = host_rounds(opts)
            assert [r.nit for r in got] == nit.tolist(), (opts, [r.nit for r in got], nit.tolist())
            assert [r.status for r in got] == status.tolist()
            for k, r in enumerate(got):
                np.testing.assert_array_equal(r.x, x[k])
                assert r.fun == fun[k] and r.lr == lrs[k], (opts, k)
            tr = prob.minimize_proximal_gradient_batched(X0, grid, return_all=True, **opts)
            for k, r in enumerate(tr):
                assert r.nit == nit[k] and len(r.allerrs) == nit[k]
                np.testing.assert_array_equal(r.x, x[k])
                if nit[k]:
                    assert r.allfuns[-1] == r.fun
    assert len(set(nit.tolist())) >= 1
