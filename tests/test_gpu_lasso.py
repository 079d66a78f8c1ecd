"""GPU parity tests of the large-n LASSO path (DESIGN.md row c): csrc/zf_lasso.cu through
zfista_b200.lasso.DenseLasso, against the reference's golden runs and the CPU oracle."""
import warnings

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu
REL = 1e-8


def _close(a, b, rel=REL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    np.testing.assert_allclose(a, b, rtol=rel, atol=rel * max(1.0, float(np.max(np.abs(b)))))


def test_gradient_matches_numpy(gpu):
    """jac_f = A.T @ (A @ x - b) * (2*scale), f = ||A x - b||^2 * scale; vector (even n_cols)
    and scalar (odd n_cols) kernels, ragged row / column counts."""
    from zfista_b200.lasso import DenseLasso

    rng = np.random.RandomState(0)
    for n_rows, n_cols in [(1, 1), (3, 1), (7, 2), (50, 200), (257, 1023), (1000, 2050),
                           (3001, 4098)]:
        A = rng.standard_normal((n_rows, n_cols))
        b = rng.standard_normal(n_rows)
        x = rng.standard_normal(n_cols)
        prob = DenseLasso(A, b, 0.1, scale=0.37)
        grad, f = prob.gradient(x)
        r = A @ x - b
        _close(grad.cpu().numpy(), A.T @ r * (2 * 0.37), rel=1e-12)
        _close(f.item(), np.linalg.norm(r) ** 2 * 0.37, rel=1e-13)
        _close(prob.jac_f(x), A.T @ r * (2 * 0.37), rel=1e-12)


def test_reference_toy_and_dataset_runs(gpu):
    """The reference's own single-objective tests (tests/test_proximal_gradient.py:66-111 toy
    problems, build_dataset() regression) and fixed-step FISTA in the cameraman notebook's
    style (lr = 1/L, decay_rate = 1, custom (a, b)): same nit, x and F within 1e-8."""
    from zfista_b200 import minimize_proximal_gradient
    from zfista_b200.lasso import DenseLasso

    d = helpers.load("lasso_single")
    for l1, nest, x_ref, fun_ref, nit_ref in d["toy_rows"]:
        prob = DenseLasso(d["toy_A"], d["toy_b"], l1, scale=1 / 6)
        res = minimize_proximal_gradient(prob.f, prob.g, prob.jac_f, prob.prox_wsum_g,
                                         d["toy_x0"], nesterov=bool(nest))
        assert res.success and res.nit == int(nit_ref)
        _close(res.x, [x_ref])
        _close(res.fun, fun_ref)
    prob = DenseLasso(d["ds_A"], d["ds_b"], float(d["ds_l1"]), scale=float(d["ds_scale"]))
    L = float(d["ds_L"])
    runs = {
        "bt_ista": dict(nesterov=False),
        "bt_fista": dict(nesterov=True),
        "fixed_fista": dict(nesterov=True, lr=1 / L, decay_rate=1),
        "fixed_fista_ab": dict(nesterov=True, lr=1 / L, decay_rate=1,
                               nesterov_ratio=(0.5, 1 / 16)),
    }
    for tag, opts in runs.items():
        res = prob.minimize_proximal_gradient(d["ds_x0"], max_iter=20000, return_all=True, **opts)
        assert res.nit == int(d[f"ds_{tag}_nit"]), tag
        _close(res.x, d[f"ds_{tag}_x"])
        _close(res.fun, d[f"ds_{tag}_fun"])
        _close(res.allerrs, d[f"ds_{tag}_allerrs"], rel=1e-6)
        _close(np.ravel(res.allfuns), np.ravel(d[f"ds_{tag}_allfuns"]))
        # without traces the fixed-step run skips the F evaluations but must land on the same x
        res2 = prob.minimize_proximal_gradient(d["ds_x0"], max_iter=20000, **opts)
        assert res2.nit == res.nit and res2.allerrs is None
        np.testing.assert_array_equal(res2.x, res.x)
        _close(res2.fun, res.fun, rel=1e-14)


@pytest.mark.parametrize("n_rows,n_cols", [(300, 120), (200, 501)])
def test_seeded_lasso_matches_oracle(gpu, n_rows, n_cols):
    from oracle import zfista_oracle as zo
    from zfista_b200.lasso import DenseLasso

    rng = np.random.RandomState(n_rows + n_cols)
    A = rng.standard_normal((n_rows, n_cols))
    w = np.zeros(n_cols)
    w[:10] = rng.standard_normal(10)
    b = A @ w + 0.01 * rng.standard_normal(n_rows)
    scale, l1 = 1 / (2 * n_rows), 0.05
    spec = zo.make_least_squares_l1(A, b, l1, scale=scale)
    prob = DenseLasso(A, b, l1, scale=scale)
    x0 = rng.standard_normal(n_cols) * 0.1
    for opts in (dict(nesterov=True), dict(nesterov=False, max_iter=300),
                 dict(nesterov=True, deprecated=True),
                 dict(nesterov=True, nesterov_ratio=(0.25, 17 / 128), lr=4.0)):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = zo.minimize_proximal_gradient(spec, x0, **opts)
            res = prob.minimize_proximal_gradient(x0, **opts)
        assert res.nit == ref["nit"] and res.success == ref["success"]
        _close(res.x, ref["x"])
        _close(res.fun, ref["fun"])


def test_backtracking_failure_and_max_iter(gpu):
    from zfista_b200.lasso import DenseLasso

    rng = np.random.RandomState(1)
    A = rng.standard_normal((40, 30)) * 100
    b = rng.standard_normal(40)
    prob = DenseLasso(A, b, 0.1)
    x0 = np.ones(30)
    res = prob.minimize_proximal_gradient(x0, lr=1e6, max_backtrack_iter=2)
    assert res.status == -1 and not res.success and res.nit == 0
    np.testing.assert_array_equal(res.x, x0)
    with pytest.warns(UserWarning, match="Maximum number of iterations"):
        res = prob.minimize_proximal_gradient(x0, max_iter=4)
    assert res.status == 0 and res.nit == 4


def test_large_a_properties(gpu):
    """A >> L2 (1.3 GB): properties that do not need a CPU pass over A.
    * the gradient is affine in x:  J(x1 + x2) + J(0) == J(x1) + J(x2);
    * f(x) >= 0 and f at the planted solution is the noise level;
    * fixed-step FISTA decreases F and recovers the planted support."""
    import torch
    from zfista_b200.lasso import DenseLasso

    n_rows, n_cols = 20000, 8192
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn(n_rows, n_cols, dtype=torch.float64, device="cuda", generator=g)
    w = torch.zeros(n_cols, dtype=torch.float64, device="cuda")
    w[:32] = torch.randn(32, dtype=torch.float64, device="cuda", generator=g) + 3.0
    b = A @ w
    prob = DenseLasso(A, b, l1_ratio=1e-3, scale=1 / (2 * n_rows))
    x1 = torch.randn(n_cols, dtype=torch.float64, device="cuda", generator=g)
    x2 = torch.randn(n_cols, dtype=torch.float64, device="cuda", generator=g)
    zero = torch.zeros_like(x1)
    j12, j0 = prob.gradient(x1 + x2)[0], prob.gradient(zero)[0]
    j1, j2 = prob.gradient(x1)[0], prob.gradient(x2)[0]
    torch.testing.assert_close(j12 + j0, j1 + j2, rtol=1e-11, atol=1e-11)
    torch.testing.assert_close(j1, (A.T @ (A @ x1 - b)) / n_rows, rtol=1e-11, atol=1e-11)
    assert prob.f(w) < 1e-20
    res = prob.minimize_proximal_gradient(zero, lr=0.4, decay_rate=1, nesterov=True,
                                          max_iter=60, return_all=True)
    F = np.array(res.allfuns)
    assert F[-1] < 1e-3 * F[0]
    x = np.asarray(res.x)
    assert np.all(np.abs(x[:32]) > 1.0) and np.max(np.abs(x[32:])) < 0.05


@pytest.mark.parametrize("env,shape,passes", [
    ({"ZF_LASSO_FUSED": "0"}, (700, 1030), 2),            # two-pass kernels
    ({}, (700, 2050), 1),                                  # single-CTA fused (default policy)
    ({"ZF_LASSO_FUSED": "1"}, (701, 3000), 1),             # single-CTA fused forced, odd rows
    ({}, (1200, 9000), 1),                                 # default: chunk ring, cluster 2
    ({}, (900, 4100), 1),                                  # default: chunk ring, no cluster
    ({"ZF_LASSO_RING": "1"}, (903, 2050), 1),              # chunk ring, no cluster, 2 chunks (ragged)
    ({"ZF_LASSO_RING": "1"}, (700, 10240), 1),             # 5 full chunks per row
    ({"ZF_LASSO_RING": "2"}, (1001, 5002), 1),             # cluster 2, ragged slices and chunks
    ({"ZF_LASSO_RING": "2"}, (640, 20000), 1),             # BASELINE configs[3] row width
    ({"ZF_LASSO_RING": "4"}, (1300, 8200), 1),             # cluster 4: few rows per cluster
    ({"ZF_LASSO_RING": "2"}, (640, 6), 1),                 # one short chunk, slices of 2 and 1 pairs
    ({"ZF_LASSO_RING": "3"}, (700, 20000), 1),             # odd cluster sizes: 3 x 4 chunks
    ({"ZF_LASSO_RING": "5"}, (650, 36002), 1),             # 5 CTAs, ragged last chunk
    ({"ZF_LASSO_RING": "6"}, (640, 40000), 1),
    ({"ZF_LASSO_RING": "7"}, (610, 30000), 1),
    ({}, (1200, 20000), 1),                                # create-time probe (4-CTA + 2-CTA split or not)
    ({"ZF_LASSO_TUNE": "v"}, (900, 36000), 1),             # probe over 5..8-CTA clusters, table printed
    ({"ZF_LASSO_TUNE": "0"}, (900, 36000), 1),             # static policy: 5-CTA clusters
    ({"ZF_LASSO_NSM": "100"}, (900, 24000), 1),            # planned for a part with fewer SMs
    ({"ZF_LASSO_NSM": "37", "ZF_LASSO_TUNE": "0"}, (640, 20000), 1),
])
def test_every_gradient_kernel_form(gpu, monkeypatch, env, shape, passes):
    """Each form of the A^T(A v - b) pass (csrc/zf_lasso.cu) forced through its environment
    switch: gradient / f against a torch fp64 reference, and a FISTA solve that must agree
    with the two-pass kernels' solve (same nit, x to 1e-9)."""
    import torch
    from zfista_b200.lasso import DenseLasso

    for k in ("ZF_LASSO_FUSED", "ZF_LASSO_RING", "ZF_LASSO_TUNE", "ZF_LASSO_NSM", "ZF_LASSO_RING_RPS"):
        monkeypatch.delenv(k, raising=False)
    rows, cols = shape
    g = torch.Generator(device="cuda").manual_seed(rows + cols)
    A = torch.randn(rows, cols, dtype=torch.float64, device="cuda", generator=g)
    w = torch.zeros(cols, dtype=torch.float64, device="cuda")
    w[:12] = 1.0
    b = A @ w + 0.01 * torch.randn(rows, dtype=torch.float64, device="cuda", generator=g)
    monkeypatch.setenv("ZF_LASSO_FUSED", "0")
    two_pass = DenseLasso(A, b, 0.02, scale=1 / (2 * rows))
    assert two_pass.hbm_passes_per_gradient() == 2
    monkeypatch.delenv("ZF_LASSO_FUSED")
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    prob = DenseLasso(A, b, 0.02, scale=1 / (2 * rows))
    assert prob.hbm_passes_per_gradient() == passes
    x = torch.randn(cols, dtype=torch.float64, device="cuda", generator=g)
    grad, f = prob.gradient(x)
    r = A @ x - b
    torch.testing.assert_close(grad, (A.T @ r) / rows, rtol=1e-12, atol=1e-13)
    torch.testing.assert_close(f[0], (r @ r) / (2 * rows), rtol=1e-13, atol=0)
    opts = dict(nesterov=True, max_iter=40)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = prob.minimize_proximal_gradient(np.zeros(cols), **opts)
        c = two_pass.minimize_proximal_gradient(np.zeros(cols), **opts)
    assert a.nit == c.nit and a.status == c.status
    np.testing.assert_allclose(a.x, c.x, rtol=0, atol=1e-9)
    np.testing.assert_allclose(a.fun, c.fun, rtol=1e-11)


@pytest.mark.parametrize("cluster,shape", [
    ("4", (901, 20000)), ("4", (1200, 8192)), ("6", (641, 36000)), ("8", (777, 40000)),
    ("3", (610, 12000)), ("2", (650, 4000)),
])
def test_ring_rows_per_exchange_step_is_bit_identical(gpu, monkeypatch, cluster, shape):
    """Two rows per exchange step (the default for row slices of <= 3 chunks: one dot-product
    reduction, one exchange and one `ready` wait per ROW PAIR; ZF_LASSO_RING_RPS=1 switches it
    off) changes the schedule of the chunk-ring kernel, not its arithmetic: gradient and f must
    equal the one-row-per-step kernel's bit for bit, also when the rows of a cluster are odd in
    number (a last step of one row)."""
    import torch
    from zfista_b200.lasso import DenseLasso

    for k in ("ZF_LASSO_FUSED", "ZF_LASSO_RING", "ZF_LASSO_TUNE", "ZF_LASSO_NSM", "ZF_LASSO_RING_RPS"):
        monkeypatch.delenv(k, raising=False)
    rows, cols = shape
    g = torch.Generator(device="cuda").manual_seed(rows * 7 + cols)
    A = torch.randn(rows, cols, dtype=torch.float64, device="cuda", generator=g)
    b = torch.randn(rows, dtype=torch.float64, device="cuda", generator=g)
    x = torch.randn(cols, dtype=torch.float64, device="cuda", generator=g)
    monkeypatch.setenv("ZF_LASSO_RING", cluster)
    out = {}
    for rps in ("1", "2"):
        monkeypatch.setenv("ZF_LASSO_RING_RPS", rps)
        prob = DenseLasso(A, b, 0.02, scale=1 / (2 * rows))
        assert prob.hbm_passes_per_gradient() == 1
        grad, f = prob.gradient(x)
        out[rps] = (grad.clone(), f.clone())
        del prob
    assert torch.equal(out["1"][0], out["2"][0])
    assert torch.equal(out["1"][1], out["2"][1])
    r = A @ x - b
    torch.testing.assert_close(out["2"][0], (A.T @ r) / rows, rtol=1e-12, atol=1e-13)


def test_reference_lasso_zero_and_return_all(gpu):
    """tests/test_proximal_gradient.py:43-68 (A = 0, b = 0: the minimiser is x = 0) and
    :221-243 (return_all fields), through both single-objective device paths."""
    import zfista_b200.problems as zp
    from zfista_b200 import minimize_proximal_gradient
    from zfista_b200.lasso import DenseLasso

    A = np.zeros((3, 1))
    b = np.zeros(3)
    x0 = np.array([0.5488135039273248])
    for prob in (zp.LeastSquaresL1(A, b, 0.1, scale=1 / 6), DenseLasso(A, b, 0.1, scale=1 / 6)):
        res = minimize_proximal_gradient(prob.f, prob.g, prob.jac_f, prob.prox_wsum_g, x0)
        res_nesterov = minimize_proximal_gradient(prob.f, prob.g, prob.jac_f, prob.prox_wsum_g, x0,
                                                  nesterov=True)
        np.testing.assert_array_almost_equal(res.x, [0], decimal=3)
        np.testing.assert_array_almost_equal(res_nesterov.x, [0], decimal=3)
        res = minimize_proximal_gradient(prob.f, prob.g, prob.jac_f, prob.prox_wsum_g, x0,
                                         return_all=True)
        assert "allvecs" in res and "allerrs" in res and "allfuns" in res
        assert len(res.allerrs) == res.nit and len(res.allfuns) == res.nit + 1


def test_baseline_configs3_full_size(gpu):
    """BASELINE configs[3] at its full size on one GPU: A 200000 x 20000 fp64 (29.8 GiB).  No CPU
    pass over A is possible in test time, so the checks are (a) an independent fp64 evaluation
    of the same gradient / f by cuBLAS (torch mv) on the same device buffer, (b) affinity in x,
    (c) the single-run chunk-ring kernel (4-CTA clusters at this width) against the multi-run
    tensor-core passes, (d) a few FISTA iterations that must agree between the two paths."""
    import torch
    from zfista_b200.lasso import DenseLasso, DenseLassoMulti

    n_rows, n_cols = 200000, 20000
    free, _ = torch.cuda.mem_get_info()
    if free < (n_rows * n_cols * 8) * 1.15:
        pytest.skip("needs ~35 GB of free HBM")
    g = torch.Generator(device="cuda").manual_seed(3)
    A = torch.empty(n_rows, n_cols, dtype=torch.float64, device="cuda")
    for r0 in range(0, n_rows, 4000):
        A[r0:r0 + 4000] = torch.randn(4000, n_cols, dtype=torch.float64, device="cuda", generator=g)
    w = torch.zeros(n_cols, dtype=torch.float64, device="cuda")
    w[:50] = 2.0
    b = A @ w + 0.1 * torch.randn(n_rows, dtype=torch.float64, device="cuda", generator=g)
    scale = 1 / (2 * n_rows)
    prob = DenseLasso(A, b, l1_ratio=1e-3, scale=scale)
    assert prob.hbm_passes_per_gradient() == 1
    x1 = torch.randn(n_cols, dtype=torch.float64, device="cuda", generator=g)
    x2 = torch.randn(n_cols, dtype=torch.float64, device="cuda", generator=g)
    zero = torch.zeros_like(x1)
    (j1, f1), (j2, _), (j0, _), (j12, _) = (prob.gradient(v) for v in (x1, x2, zero, x1 + x2))
    r1 = torch.mv(A, x1) - b
    torch.testing.assert_close(j1, torch.mv(A.T, r1) * (2 * scale), rtol=1e-11, atol=1e-11)
    torch.testing.assert_close(f1[0], (r1 @ r1) * scale, rtol=1e-12, atol=0)
    torch.testing.assert_close(j12 + j0, j1 + j2, rtol=1e-11, atol=1e-11)
    multi = DenseLassoMulti(A, b, 1e-3, 3, scale=scale)
    G, F = multi.gradient(torch.stack([x1, x2, zero]))
    torch.testing.assert_close(G[0], j1, rtol=1e-11, atol=1e-11)
    torch.testing.assert_close(G[1], j2, rtol=1e-11, atol=1e-11)
    torch.testing.assert_close(G[2], j0, rtol=1e-11, atol=1e-11)
    torch.testing.assert_close(F[0], f1[0], rtol=1e-12, atol=0)
    opts = dict(lr=0.4, decay_rate=1, nesterov=True, max_iter=6, tol=0.0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = prob.minimize_proximal_gradient(zero, **opts)
        m = multi.minimize_proximal_gradient_batched(zero, [(0, 0.25), (0.5, 1 / 16), (0, 0.25)], **opts)
    assert a.nit == m[0].nit == 6
    np.testing.assert_allclose(m[0].x, a.x, rtol=0, atol=1e-10)
    np.testing.assert_allclose(m[2].x, m[0].x, rtol=0, atol=0)      # same run twice: bit-identical
    np.testing.assert_allclose(m[0].fun, a.fun, rtol=1e-11)
    del A
    torch.cuda.empty_cache()


@pytest.mark.parametrize("n_rows,n_cols", [(300, 120), (700, 4100)])
def test_device_decided_loop_equals_host_decided_loop(gpu, n_rows, n_cols):
    """zf_lasso_solve runs the device-decided loop (scalars and decisions in device memory, CUDA
    graph of 32 trials, no host sync per trial); the split begin / grad / step / finish entry
    points still take every decision on the host.  Both must produce the SAME bits -- nit, status,
    x, F and the traces -- with a line search (retries in the first iterations), a fixed step, a
    line-search failure and max_iter stops, and across repeated solves on one handle (the graph
    is captured once and replayed)."""
    import ctypes as C

    import torch

    from zfista_b200 import _lib
    from zfista_b200.lasso import DenseLasso
    from zfista_b200.proximal_gradient import _make_options

    rng = np.random.RandomState(n_rows)
    A = rng.standard_normal((n_rows, n_cols))
    w = np.zeros(n_cols)
    w[:10] = rng.standard_normal(10)
    b = A @ w + 0.01 * rng.standard_normal(n_rows)
    scale, l1 = 1 / (2 * n_rows), 0.05
    prob = DenseLasso(A, b, l1, scale=scale)
    L = _lib.lib()
    x0 = rng.standard_normal(n_cols) * 0.1
    lip = 2 * scale * np.linalg.norm(A, 2) ** 2

    def host_loop(opts):
        o = _make_options(opts.get("lr", 1), opts.get("tol", 1e-5), 1e-12,
                          opts.get("max_iter", 1000000), 100000,
                          opts.get("max_backtrack_iter", 100), False, opts.get("decay_rate", 0.5),
                          opts.get("nesterov", False), opts.get("nesterov_ratio", (0, 0.25)),
                          opts.get("deprecated", False), "reference", 0)
        x0d = torch.from_numpy(x0).cuda()
        xd = torch.empty_like(x0d)
        fun, nit, status, nxt = C.c_double(), C.c_int64(), C.c_int32(), C.c_int32(0)
        _lib.check(L.zf_lasso_begin(prob._h, C.byref(o), C.c_void_p(x0d.data_ptr())))
        _lib.check(L.zf_lasso_step(prob._h, C.byref(nxt)))
        while nxt.value != 2:
            _lib.check(L.zf_lasso_grad(prob._h, nxt.value))
            _lib.check(L.zf_lasso_step(prob._h, C.byref(nxt)))
        _lib.check(L.zf_lasso_finish(prob._h, C.c_void_p(xd.data_ptr()), C.byref(fun),
                                     C.byref(nit), C.byref(status)))
        return xd.cpu().numpy(), fun.value, nit.value, status.value

    cases = [dict(nesterov=True), dict(nesterov=False, max_iter=40),
             dict(nesterov=True, lr=1 / lip, decay_rate=1, max_iter=300),
             dict(nesterov=True, lr=1 / lip, decay_rate=1, max_iter=33, tol=0.0),
             dict(nesterov=True, deprecated=True, lr=8.0),
             dict(nesterov=True, lr=1e9, max_backtrack_iter=3),
             dict(nesterov=True, nesterov_ratio=(0.5, 1 / 16))]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for opts in cases + cases[:3]:
            res = prob.minimize_proximal_gradient(x0, **opts)
            x, fun, nit, status = host_loop(opts)
            assert (res.nit, res.status) == (nit, status), (opts, res.nit, nit, res.status, status)
            np.testing.assert_array_equal(res.x, x)
            assert res.fun == fun, (opts, res.fun, fun)
            tr = prob.minimize_proximal_gradient(x0, return_all=True, **opts)
            assert tr.nit == nit and len(tr.allerrs) == nit and len(tr.allfuns) == nit + 1
            np.testing.assert_array_equal(tr.x, x)
            if nit:
                assert tr.allfuns[-1] == tr.fun


def test_return_all_records_the_iterates(gpu):
    """return_all=True on the large-n path: allvecs = [x^0, ..., x^nit] as the reference records
    them (proximal_gradient.py:471, 522), with line-search retries in the first iterations (a
    rejected candidate must not be recorded); a byte budget truncates; "funs" leaves them out."""
    from oracle import zfista_oracle as zo
    from zfista_b200.lasso import DenseLasso

    rng = np.random.RandomState(7)
    n_rows, n_cols = 250, 140
    A = rng.standard_normal((n_rows, n_cols))
    w = np.zeros(n_cols)
    w[:9] = rng.standard_normal(9)
    b = A @ w + 0.01 * rng.standard_normal(n_rows)
    scale, l1 = 1 / (2 * n_rows), 0.04
    spec = zo.make_least_squares_l1(A, b, l1, scale=scale)
    prob = DenseLasso(A, b, l1, scale=scale)
    x0 = rng.standard_normal(n_cols) * 0.1
    for opts in (dict(nesterov=True, lr=16.0), dict(nesterov=False, max_iter=60)):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = zo.minimize_proximal_gradient(spec, x0, return_all=True, **opts)
            res = prob.minimize_proximal_gradient(x0, return_all=True, **opts)
        assert res.nit == ref["nit"] and len(res.allvecs) == res.nit + 1
        assert not res.allvecs_truncated
        _close(np.array(res.allvecs), np.array(ref["allvecs"]))
        np.testing.assert_array_equal(res.allvecs[0], x0)
        np.testing.assert_array_equal(res.allvecs[-1], res.x)
        _close(np.ravel(res.allfuns), np.ravel(ref["allfuns"]))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cut = prob.minimize_proximal_gradient(x0, return_all=True, nesterov=True,
                                              allvecs_bytes=5 * n_cols * 8)
        funs = prob.minimize_proximal_gradient(x0, return_all="funs", nesterov=True)
    assert len(cut.allvecs) == 5 and cut.allvecs_truncated and len(cut.allerrs) == cut.nit
    full = prob.minimize_proximal_gradient(x0, return_all=True, nesterov=True)
    np.testing.assert_array_equal(np.array(cut.allvecs), np.array(full.allvecs[:5]))
    assert funs.allvecs is None and len(funs.allfuns) == funs.nit + 1
