"""GPU parity tests of the batched proximal-gradient path (rows a, b of DESIGN.md).

Everything here calls the CUDA kernels through the C ABI (zfista_b200._lib) and
compares with
  * the golden fixtures produced by the unmodified reference (tests/golden/), and
  * the CPU oracle (oracle/zfista_oracle.py) on fresh seeded inputs.

Tolerances (BASELINE.json north_star): the same iteration count per start, final x and
F(x) within 1e-8 relative.  The device functors themselves agree with the reference to
a few ulp (summation order differs: warp-strided vs numpy pairwise).
"""
import warnings

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu

FIX = helpers.fixture_problems()
REL = 1e-8


def _rel_close(a, b, rel=REL, floor=1e-8):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = np.maximum(np.max(np.abs(b)), 1.0)
    np.testing.assert_allclose(a, b, rtol=rel, atol=floor * scale)


# ------------------------------------------------------------------ device functors
@pytest.mark.parametrize("pname", sorted(FIX))
def test_device_functors_match_reference(gpu, pname):
    """Problem.f / g / jac_f / prox_wsum_g on device == reference values."""
    d = helpers.load("problem_eval")
    cls, kw = FIX[pname]
    prob = helpers.device_problem(cls, kw)
    X, W = d[pname + "__X"], d[pname + "__W"]
    for k in range(len(X)):
        np.testing.assert_allclose(prob.f(X[k]), d[pname + "__f"][k], rtol=1e-13, atol=1e-300)
        np.testing.assert_allclose(prob.g(X[k]), d[pname + "__g"][k], rtol=1e-13, atol=1e-300)
        np.testing.assert_allclose(prob.jac_f(X[k]), d[pname + "__jac"][k], rtol=1e-13,
                                   atol=1e-300)
        np.testing.assert_allclose(prob.prox_wsum_g(W[k], X[k]), d[pname + "__prox"][k],
                                   rtol=1e-14, atol=1e-15)


def test_reference_unit_vectors_on_device(gpu):
    """Known answers of the reference's tests/test_problems.py, through the GPU."""
    import zfista_b200.problems as zp

    x = np.array([1, 2, 3, 4, 5], dtype=np.float64)
    jos1 = zp.JOS1()
    np.testing.assert_almost_equal(jos1.f(x), [11, 3])
    np.testing.assert_almost_equal(
        jos1.jac_f(x), [[2 / 5, 4 / 5, 6 / 5, 8 / 5, 2], [-2 / 5, 0, 2 / 5, 4 / 5, 6 / 5]])
    jl1 = zp.JOS1(l1_ratios=[0.2, 0.1], l1_shifts=[0, 1])
    np.testing.assert_almost_equal(jl1.g(x), [3, 1])
    np.testing.assert_almost_equal(
        jl1.prox_wsum_g(np.array([0.5, 0.5]), np.array([3, 4, 5, 6, 7.0])),
        [2.85, 3.85, 4.85, 5.85, 6.85])
    sd = zp.SD()
    xs = np.array([1, np.sqrt(2), np.sqrt(2), 1])
    np.testing.assert_almost_equal(sd.f(xs), [7, 8])
    np.testing.assert_almost_equal(sd.g(xs), [0, 0])
    fds = zp.FDS(n_features=5)
    np.testing.assert_almost_equal(fds.f(x), [0.0, 75.0855369, 0.1183459])
    fc = zp.FDS(n_features=5, bounds=(0, np.inf))
    assert np.all(np.isinf(fc.g(-np.ones(5))))
    np.testing.assert_almost_equal(
        fc.prox_wsum_g(np.ones(3) / 3, np.array([-3, -1, 0, 1, 3.0])), [0, 0, 0, 1, 3])
    with pytest.raises(ValueError):
        jos1.f(np.ones(4))


# ------------------------------------------------------------------ subproblem (row b)
@pytest.mark.parametrize("pname", ["JOS1_n5", "JOS1_n5_l1", "JOS1_n50_l1", "SD", "ZDT1_n50",
                                   "TOI4_l1"])
def test_two_objective_subproblem_matches_reference(gpu, pname):
    """m = 2: the device's bounded Brent lands on the reference's weight."""
    from zfista_b200 import solve_subproblems

    d = helpers.load("subproblem")
    cls, kw = FIX[pname]
    prob = helpers.device_problem(cls, kw)
    Y, XO, LR = d[pname + "__y"], d[pname + "__xold"], d[pname + "__lr"]
    dep = (np.arange(len(Y)) % 4 == 3)
    x, fun, w = solve_subproblems(prob, Y, XO, LR, deprecated=dep, tol_internal=1e-11)
    # Brent's own resolution is sqrt(eps)*|w| + tol/3: weights agree to that, the dual
    # value (flat at the optimum) much better
    np.testing.assert_allclose(w, d[pname + "__w"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(x, d[pname + "__x"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(fun, d[pname + "__fun"], rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("pname", ["TRIDIA", "TRIDIA_l1", "LFR1_n30", "FDS_n10", "FDS_n10_l1",
                                   "FDS_n10_box", "FDS_n100_l1"])
def test_multi_objective_dual_not_worse_than_reference(gpu, pname):
    """m >= 3: the device solves the dual QP exactly; its dual value must be >= what the
    reference's trust-constr reached, and equal to the CPU model of the same solver."""
    from oracle import dual_model as dm
    from oracle import zfista_oracle as zo
    from zfista_b200 import solve_subproblems

    d = helpers.load("subproblem")
    cls, kw = FIX[pname]
    prob = helpers.device_problem(cls, kw)
    spec = helpers.oracle_spec(cls, kw)
    Y, XO, LR, Wref = (d[pname + "__" + k] for k in ("y", "xold", "lr", "w"))
    dep = (np.arange(len(Y)) % 4 == 3)
    x, fun, w = solve_subproblems(prob, Y, XO, LR, deprecated=dep, tol_internal=1e-11)
    assert np.all(w >= 0) and np.allclose(w.sum(axis=1), 1, atol=1e-13)
    for k in range(len(Y)):
        fy = zo.f(spec, Y[k])
        Fp = zo.f(spec, XO[k]) + zo.g(spec, XO[k])
        J = zo.jac_f(spec, Y[k])
        c = np.zeros_like(fy) if dep[k] else fy - Fp
        args = (Y[k], J, LR[k], c, spec.l1_ratios, spec.l1_shifts, spec.lower, spec.upper,
                lambda p: zo.g(spec, p))
        D_ref = dm.dual_eval(Wref[k], *args)[0]
        D_dev, G, _, _ = dm.dual_eval(w[k], *args)
        scale = abs(D_dev) + np.max(np.abs(G))
        assert D_dev >= D_ref - 1e-12 * scale
        w_cpu, D_cpu, p_cpu, _ = dm.simplex_newton(*args)
        assert abs(D_dev - D_cpu) <= 1e-11 * scale
        assert abs(fun[k] - D_cpu) <= 1e-10 * scale


# ------------------------------------------------------------------ full solves (row a)
def _two_objective_cases():
    return [c for c in helpers.golden_cases()
            if FIX.get(c.split("__")[0], ("",))[0] in ("JOS1", "SD", "ZDT1", "TOI4")]


# Cases on which the reference is NOT reproducible to 1e-8 even against itself (a 1-ulp
# perturbation of its dual function moves nit by ~10 % and x by up to 1e-3): L1 terms and
# TOI4's flat direction.  See helpers.oracle_noise_envelope and DESIGN.md "Parity".
def _reference_is_rounding_sensitive(case):
    pname = case.split("__")[0]
    return "_l1" in pname or pname.startswith("TOI4")


@pytest.mark.parametrize("case", _two_objective_cases())
def test_batched_solve_matches_reference(gpu, case):
    """Every start of a golden case in ONE launch.  Well-conditioned cases: the same nit per
    start, x and F within 1e-8 relative.  Rounding-sensitive cases: within the envelope the
    reference shows against itself under a 1-ulp perturbation."""
    d = helpers.load(case)
    cls, kw = str(d["problem"]), helpers.case_kwargs(d)
    prob = helpers.device_problem(cls, kw)
    opts = helpers.case_options(d)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        br = prob.minimize_proximal_gradient_batched(d["x0"], return_all=True, **opts)
    np.testing.assert_array_equal(br.success, d["success"])
    if not _reference_is_rounding_sensitive(case):
        np.testing.assert_array_equal(br.nit, d["nit"])
        _rel_close(br.x, d["x"])
        _rel_close(br.fun, d["fun"])
        n0 = int(d["nit"][0])
        _rel_close(br.allerrs[0, :n0], d["allerrs0"], rel=1e-6, floor=1e-9)
        _rel_close(br.allfuns[0, :n0 + 1], d["allfuns0"])
        return
    # the reference's own envelope: its dual value perturbed by ONE ulp, 8 seeds.  Measured on
    # B200 (profiles/r02_envelope_probe.py, all 15 such cases): the device's iteration counts are
    # inside 1.0x that envelope on every case (e.g. JOS1 n=50 +L1 FISTA: 16 against 30; TOI4 +L1
    # FISTA: 8 against 15; 7 cases 0 against 0), x inside 1.33x (1.29e-3 against 9.7e-4 on
    # JOS1_n50_l1nb FISTA, <= 1.0x elsewhere), F inside 1.07x.  Asserted: 1x, 1.5x, 1.5x.
    env = helpers.oracle_noise_envelope(helpers.oracle_spec(cls, kw), d["x0"], d["x"], d["fun"],
                                        d["nit"], opts,
                                        n_starts=16 if prob.n_features <= 10 else 4,
                                        seeds=tuple(range(8)))
    dnit = np.abs(br.nit - d["nit"])
    dx = np.max(np.abs(br.x - d["x"]))
    dF = np.max(np.abs(br.fun - d["fun"]) / np.maximum(1.0, np.abs(d["fun"])))
    assert dnit.max() <= env["dnit"], (dnit, env)
    assert dx <= 1.5 * env["dx"] + 1e-8, (dx, env)
    assert dF <= 1.5 * env["dF"] + 1e-8, (dF, env)
    # the first iterations (before the noise is amplified) still agree tightly
    k = min(5, int(min(br.nit[0], d["nit"][0])))
    _rel_close(br.allerrs[0, :k], d["allerrs0"][:k], rel=1e-5, floor=1e-8)
    _rel_close(br.allfuns[0, :k + 1], d["allfuns0"][:k + 1], rel=1e-7)


def _multi_objective_cases():
    return [c for c in helpers.golden_cases()
            if FIX.get(c.split("__")[0], ("",))[0] in ("TRIDIA", "FDS", "LinearFunctionRank1")]


@pytest.mark.parametrize("case", _multi_objective_cases())
def test_multi_objective_solve_tracks_reference(gpu, case):
    """m >= 3.  The reference's inner solver here is scipy trust-constr, which is neither
    reproducible step by step nor exact (it reaches the dual optimum to ~1e-6 when it
    converges within max_iter_internal, and the fixtures bound it at 1000 inner / 40 outer
    iterations because single subproblems otherwise take minutes).  The device solves the
    dual exactly, so the claim is *tracking*: F(x^k) follows the reference's trajectory to
    trust-constr's own accuracy, and where the reference genuinely converged the device
    converges in a comparable number of iterations to a point with the same F."""
    d = helpers.load(case)
    prob = helpers.device_problem(str(d["problem"]), helpers.case_kwargs(d))
    opts = helpers.case_options(d)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        br = prob.minimize_proximal_gradient_batched(d["x0"], return_all=True, **opts)
    assert np.all(br.status >= 0)                      # the exact dual never breaks the line search
    k = int(min(br.nit[0], d["nit"][0]))
    F_dev, F_ref = br.allfuns[0, :k + 1], d["allfuns0"][:k + 1]
    rel = np.abs(F_dev - F_ref) / np.maximum(1.0, np.abs(F_ref))
    assert rel.max() < 5e-3, rel.max()
    assert rel[:3].max() < 1e-4                        # before trust-constr's error accumulates
    for i in range(len(d["nit"])):
        ref_nit, ref_ok = int(d["nit"][i]), bool(d["success"][i])
        if ref_ok and ref_nit >= 3:
            # (ref_nit < 3 "successes" are step-size collapses of the inexact dual, e.g.
            # FDS n = 100 stops after 2 iterations at F_1 = 1.7e7 -- not a minimum)
            slack = max(3, 0.3 * ref_nit)
            if br.status[i] == 0:      # the fixture's outer cap cut the device off just short
                assert ref_nit + slack >= opts["max_iter"], (br.nit[i], ref_nit)
                continue
            assert br.status[i] == 1
            assert abs(int(br.nit[i]) - ref_nit) <= slack, (br.nit[i], ref_nit)
            # both stop at Pareto-stationary points of the same front; which one depends on
            # the path, so F agrees to the path's accuracy, relative to the largest objective
            scale = max(1.0, float(np.max(np.abs(d["fun"][i]))))
            np.testing.assert_allclose(br.fun[i], d["fun"][i], rtol=0, atol=5e-3 * scale)


def test_single_start_api_matches_reference_fields(gpu):
    """minimize_proximal_gradient(f, g, jac_f, prox_wsum_g, x0, ...) drop-in call."""
    from zfista_b200 import minimize_proximal_gradient
    import zfista_b200.problems as zp

    d = helpers.load("JOS1_n5__fista")
    prob = zp.JOS1(n_features=5)
    opts = helpers.case_options(d)
    res = minimize_proximal_gradient(prob.f, prob.g, prob.jac_f, prob.prox_wsum_g,
                                     d["x0"][0], return_all=True, **opts)
    assert res.success and res.status == 1 and res.nit == int(d["nit"][0])
    assert res.message == "Optimization terminated successfully"
    for key in ("x", "fun", "nit", "nfev", "success", "time", "allvecs", "allfuns", "allerrs",
                "x0", "tol", "tol_internal", "nesterov", "nesterov_ratio"):
        assert key in res
    _rel_close(res.x, d["x"][0])
    _rel_close(res.fun, d["fun"][0])
    assert len(res.allvecs) == res.nit + 1 and len(res.allerrs) == res.nit
    # method form
    res2 = prob.minimize_proximal_gradient(d["x0"][0], **opts)
    assert res2.nit == res.nit and res2.allvecs is None
    with pytest.raises(TypeError):
        minimize_proximal_gradient(lambda x: x, prob.g, prob.jac_f, prob.prox_wsum_g,
                                   d["x0"][0])


def test_max_iter_and_warning(gpu):
    import zfista_b200.problems as zp

    prob = zp.JOS1(n_features=50)
    x0 = np.linspace(-2, 4, 50)
    with pytest.warns(UserWarning, match="Maximum number of iterations"):
        res = prob.minimize_proximal_gradient(x0, max_iter=3)
    assert not res.success and res.status == 0 and res.nit == 3


def test_per_start_momentum_grid_matches_per_pair_runs(gpu):
    """(a, b) sweep: one launch with a per-start (a, b) table == one launch per pair."""
    import zfista_b200.problems as zp

    prob = zp.JOS1(n_features=50, l1_ratios=(1 / 50, 1 / 100), l1_shifts=(0, 1))
    rng = np.random.RandomState(3)
    X0 = rng.uniform(-2, 4, size=(6, 50))
    grid = np.array(helpers.AB_GRID)
    X0g = np.repeat(X0, len(grid), axis=0)
    ABg = np.tile(grid, (len(X0), 1))
    all_in_one = prob.minimize_proximal_gradient_batched(X0g, nesterov=True,
                                                         nesterov_ratio=ABg, tol_internal=1e-11)
    for gi, ab in enumerate(grid):
        one = prob.minimize_proximal_gradient_batched(X0, nesterov=True,
                                                      nesterov_ratio=tuple(ab), tol_internal=1e-11)
        sel = np.arange(len(X0)) * len(grid) + gi
        np.testing.assert_array_equal(all_in_one.nit[sel], one.nit)
        np.testing.assert_array_equal(all_in_one.x[sel], one.x)


@pytest.mark.parametrize("cls,kw,lo,hi", [
    ("JOS1", dict(n_features=20), -2, 4),
    ("JOS1", dict(n_features=100, l1_ratios=(0.01, 0.02), l1_shifts=(0, 1)), -2, 4),
    ("ZDT1", dict(n_features=100), 0, 0.01),
    ("TOI4", dict(bounds=(-1.0, 3.0)), -1, 3),
])
@pytest.mark.parametrize("algo", ["ista", "fista"])
def test_fresh_starts_match_oracle(gpu, cls, kw, lo, hi, algo):
    """Seeded random starts not in the fixtures: CUDA path vs the CPU oracle."""
    from oracle import zfista_oracle as zo

    prob = helpers.device_problem(cls, kw)
    spec = helpers.oracle_spec(cls, kw)
    rng = np.random.RandomState(sum(map(ord, cls + algo)))
    X0 = rng.uniform(lo, hi, size=(5, prob.n_features))
    opts = dict(nesterov=(algo == "fista"), tol_internal=1e-11)
    br = prob.minimize_proximal_gradient_batched(X0, **opts)
    assert np.all(br.status == 1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = [zo.minimize_proximal_gradient(spec, X0[i], **opts) for i in range(len(X0))]
    x_ref = np.stack([r["x"] for r in ref])
    f_ref = np.stack([r["fun"] for r in ref])
    n_ref = np.array([r["nit"] for r in ref])
    if "l1_ratios" not in kw and cls != "TOI4":
        np.testing.assert_array_equal(br.nit, n_ref)
        _rel_close(br.x, x_ref)
        _rel_close(br.fun, f_ref)
    else:   # rounding-sensitive: within the reference's own envelope
        env = helpers.oracle_noise_envelope(spec, X0, x_ref, f_ref, n_ref, opts, n_starts=3,
                                            seeds=(0, 1))
        assert np.abs(br.nit - n_ref).max() <= 2 * env["dnit"] + 2
        assert np.max(np.abs(br.x - x_ref)) <= 4 * env["dx"] + 1e-6
        assert np.max(np.abs(br.fun - f_ref) / np.maximum(1, np.abs(f_ref))) <= 4 * env["dF"] + 1e-7


def test_single_objective_lasso_matches_reference(gpu):
    """LeastSquaresL1 functor (m = 1): toy problems of the reference's
    tests/test_proximal_gradient.py and its build_dataset() regression."""
    import zfista_b200.problems as zp

    d = helpers.load("lasso_single")
    for l1, nest, x_ref, fun_ref, nit_ref in d["toy_rows"]:
        prob = zp.LeastSquaresL1(d["toy_A"], d["toy_b"], l1, scale=1 / 6)
        res = prob.minimize_proximal_gradient(d["toy_x0"], nesterov=bool(nest))
        assert res.nit == int(nit_ref)
        _rel_close(res.x, [x_ref])
        _rel_close(res.fun, fun_ref)
    prob = zp.LeastSquaresL1(d["ds_A"], d["ds_b"], float(d["ds_l1"]), scale=float(d["ds_scale"]))
    L = float(d["ds_L"])
    runs = {
        "bt_ista": dict(nesterov=False),
        "bt_fista": dict(nesterov=True),
        "fixed_fista": dict(nesterov=True, lr=1 / L, decay_rate=1),
        "fixed_fista_ab": dict(nesterov=True, lr=1 / L, decay_rate=1,
                               nesterov_ratio=(0.5, 1 / 16)),
    }
    for tag, opts in runs.items():
        res = prob.minimize_proximal_gradient(d["ds_x0"], max_iter=20000, return_all=True,
                                              **opts)
        assert res.nit == int(d[f"ds_{tag}_nit"]), tag
        _rel_close(res.x, d[f"ds_{tag}_x"])
        _rel_close(res.fun, d[f"ds_{tag}_fun"])


def test_multiobjective_toy_lasso(gpu):
    """Replicated objectives (tests/test_proximal_gradient.py:113-213): expected optima."""
    import zfista_b200.problems as zp

    A = np.array([[-1.0], [0.0], [1.0]])
    b = np.array([-1.0, 0.0, 1.0])
    expected = {1e-8: 1.0, 0.1: 0.85, 0.5: 0.25, 1.0: 0.0}
    for m in (2, 3):
        for l1, xs in expected.items():
            prob = zp.LeastSquaresL1(A, b, l1, scale=1 / 6, n_objectives=m)
            for nest in (False, True):
                res = prob.minimize_proximal_gradient(np.array([0.3745401188473625]),
                                                      nesterov=nest)
                assert res.success
                np.testing.assert_almost_equal(res.x, [xs], decimal=3)


def test_scale_properties_1024_starts(gpu):
    """BASELINE-size batch (1024 starts, FDS n = 100 with L1): size-independent
    properties -- every start converges, the result does not depend on where a start
    sits in the batch (permutation invariance) or on the batch it is launched with, and
    the solution is Pareto-stationary (one more proximal step barely moves it)."""
    import zfista_b200.problems as zp

    n = 100
    prob = zp.FDS(n_features=n, l1_ratios=(np.arange(3) + 1) / n, l1_shifts=np.arange(3))
    rng = np.random.RandomState(0)
    X0 = rng.uniform(-2, 2, size=(1024, n))
    opts = dict(nesterov=True, tol_internal=1e-11, max_iter=100000)
    br = prob.minimize_proximal_gradient_batched(X0, **opts)
    assert np.all(br.status == 1)
    perm = rng.permutation(1024)
    br2 = prob.minimize_proximal_gradient_batched(X0[perm], **opts)
    np.testing.assert_array_equal(br2.nit, br.nit[perm])
    np.testing.assert_array_equal(br2.x, br.x[perm])
    br3 = prob.minimize_proximal_gradient_batched(X0[:7], **opts)
    np.testing.assert_array_equal(br3.x, br.x[:7])
    # stationarity: one more proximal step from x* (with the step size the solve ended on)
    # moves it by no more than a small multiple of tol
    from zfista_b200 import solve_subproblems

    xs, _, w = solve_subproblems(prob, br.x[:256], br.x[:256], br.lr[:256], tol_internal=1e-11)
    assert np.max(np.abs(xs - br.x[:256])) < 1e-4
    assert np.all(w >= 0) and np.allclose(w.sum(axis=1), 1)


def test_empty_batch_and_errors(gpu):
    import zfista_b200.problems as zp
    from zfista_b200 import _lib

    prob = zp.JOS1()
    br = prob.minimize_proximal_gradient_batched(np.zeros((0, 5)))
    assert len(br) == 0
    with pytest.raises(ValueError):
        prob.minimize_proximal_gradient_batched(np.zeros((3, 4)))
    with pytest.raises(_lib.ZfError):
        prob.minimize_proximal_gradient_batched(np.zeros((1, 5)), lr=-1.0)


def test_device_entry_point_matches_host_entry_point(gpu):
    """zf_solve_batched_device (device pointers + stream, what bench.py times) gives exactly
    what the host entry point gives."""
    import ctypes as C

    import torch

    from zfista_b200 import _lib
    import zfista_b200.problems as zp
    from zfista_b200.proximal_gradient import _make_options

    prob = zp.JOS1(n_features=50, l1_ratios=(0.02, 0.01), l1_shifts=(0, 1))
    rng = np.random.RandomState(9)
    X0 = rng.uniform(-2, 4, size=(300, 50))
    host = prob.minimize_proximal_gradient_batched(X0, nesterov=True, tol_internal=1e-11)
    dev = torch.device("cuda", 0)
    x0 = torch.from_numpy(X0).to(dev)
    x = torch.empty(300, 50, dtype=torch.float64, device=dev)
    fun = torch.empty(300, 2, dtype=torch.float64, device=dev)
    nit = torch.empty(300, dtype=torch.int64, device=dev)
    status = torch.empty(300, dtype=torch.int32, device=dev)
    res = _lib.ZfResult()
    res.x, res.fun, res.nit, res.status = x.data_ptr(), fun.data_ptr(), nit.data_ptr(), status.data_ptr()
    opts = _make_options(1, 1e-5, 1e-11, 1000000, 100000, 100, False, 0.5, True, (0, 0.25), False,
                         "reference", 0)
    desc, keep = prob.descriptor()
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        _lib.check(_lib.lib().zf_solve_batched_device(
            C.byref(desc), C.byref(opts), 300, C.c_void_p(x0.data_ptr()), None, C.byref(res),
            C.c_void_p(stream.cuda_stream)))
    stream.synchronize()
    np.testing.assert_array_equal(nit.cpu().numpy(), host.nit)
    np.testing.assert_array_equal(x.cpu().numpy(), host.x)
    np.testing.assert_array_equal(fun.cpu().numpy(), host.fun)
    assert int((status == 1).sum()) == 300


def test_option_coverage(gpu):
    """warm_start, dual_solver="newton" for two objectives, array bounds, 4 objectives, and
    more than 2368 starts (the 4-warps-per-CTA launch shape)."""
    import zfista_b200.problems as zp

    rng = np.random.RandomState(4)
    prob = zp.JOS1(n_features=20)
    X0 = rng.uniform(-2, 4, size=(40, 20))
    ref = prob.minimize_proximal_gradient_batched(X0, nesterov=True, tol_internal=1e-11)
    nwt = prob.minimize_proximal_gradient_batched(X0, nesterov=True, tol_internal=1e-11,
                                                  dual_solver="newton")
    # exact dual vs Brent: same iteration counts on this well-conditioned problem, x to 1e-6
    np.testing.assert_array_equal(nwt.nit, ref.nit)
    np.testing.assert_allclose(nwt.x, ref.x, rtol=0, atol=1e-6)
    assert nwt.n_dual.sum() < ref.n_dual.sum() / 5        # and far fewer dual evaluations
    ws = prob.minimize_proximal_gradient_batched(X0, nesterov=True, tol_internal=1e-11,
                                                 warm_start=True)
    np.testing.assert_array_equal(ws.nit, ref.nit)        # Brent has no initial guess to warm
    # array bounds == the same scalar bounds
    lo, hi = np.full(20, -0.5), np.full(20, 1.5)
    pa = zp.JOS1(n_features=20, bounds=(lo, hi)).minimize_proximal_gradient_batched(X0, nesterov=True)
    ps = zp.JOS1(n_features=20, bounds=(-0.5, 1.5)).minimize_proximal_gradient_batched(X0, nesterov=True)
    np.testing.assert_array_equal(pa.x, ps.x)
    assert pa.x.min() >= -0.5 and pa.x.max() <= 1.5
    # four objectives
    lfr = zp.LinearFunctionRank1(n_features=30)
    r4 = lfr.minimize_proximal_gradient_batched(rng.uniform(-1, 1, size=(16, 30)), nesterov=True,
                                                tol_internal=1e-11)
    assert np.all(r4.status == 1) and r4.fun.shape == (16, 4)
    # big batch: every start is independent of the launch shape
    Xb = rng.uniform(-2, 4, size=(3000, 20))
    big = prob.minimize_proximal_gradient_batched(Xb, nesterov=True, tol_internal=1e-11)
    small = prob.minimize_proximal_gradient_batched(Xb[1000:1040], nesterov=True, tol_internal=1e-11)
    np.testing.assert_array_equal(big.x[1000:1040], small.x)
    np.testing.assert_array_equal(big.nit[1000:1040], small.nit)


def test_largest_benchmark_dimension_matches_oracle(gpu):
    """JOS1 with n_features = 1000, the largest size in benchmarks/benchmark.py:418 (40 KB of
    shared memory per start): same nit, x and F within 1e-8 of the CPU oracle."""
    from oracle import zfista_oracle as zo
    import zfista_b200.problems as zp

    prob = zp.JOS1(n_features=1000)
    spec = zo.make_spec("JOS1", n_features=1000)
    rng = np.random.RandomState(12)
    X0 = rng.uniform(-2, 4, size=(3, 1000))
    opts = dict(nesterov=True, tol_internal=1e-11)
    br = prob.minimize_proximal_gradient_batched(X0, **opts)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(2):
            r = zo.minimize_proximal_gradient(spec, X0[i], **opts)
            assert br.nit[i] == r["nit"]
            _rel_close(br.x[i], r["x"])
            _rel_close(br.fun[i], r["fun"])
    with pytest.raises(Exception, match="shared memory"):
        zp.JOS1(n_features=20000).minimize_proximal_gradient_batched(np.zeros((1, 20000)))


def test_return_all_is_ragged_and_survives_the_benchmark_shape(gpu):
    """benchmarks/benchmark.py:320-372 asks return_all=True for every one of its starts.  The
    traces are ragged (exactly nit_i entries per start): 1000 starts x n = 1000 with ~155
    iterations each keep 1.25 GB of iterates -- the reference keeps the same -- where the dense
    n_starts x capacity x n layout of round 1 asked for 32 GB on the host AND the device; they
    cross the device in groups of starts of at most 1 GiB.  The traces equal the single-start
    call's."""
    import zfista_b200.problems as zp

    n, S = 1000, 1000
    prob = zp.JOS1(n_features=n)
    X0 = np.random.RandomState(12).uniform(-2, 4, size=(S, n))
    br = prob.minimize_proximal_gradient_batched(X0, nesterov=True, tol_internal=1e-11,
                                                 return_all=True)
    assert np.all(br.status == 1) and not br.trace_truncated
    total = int(br.nit.sum())
    assert br.allerrs.flat.shape == (total,)
    assert br.allfuns.flat.shape == (total + S, 2)
    assert br.allvecs.flat.shape == (total + S, n)
    assert br.allvecs.nbytes == (total + S) * n * 8 < 2 << 30
    for i in (0, 17, S - 1):
        k = int(br.nit[i])
        assert br.allerrs[i].shape == (k,) and br.allvecs[i].shape == (k + 1, n)
        np.testing.assert_array_equal(br.allvecs[i][0], X0[i])
        np.testing.assert_array_equal(br.allvecs[i][-1], br.x[i])
        np.testing.assert_array_equal(br.allfuns[i][-1], br.fun[i])
        assert br.allerrs[i][-1] == br.err[i] and br.allerrs[i][-1] < 1e-5
        one = prob.minimize_proximal_gradient(X0[i], nesterov=True, tol_internal=1e-11,
                                              return_all=True)
        assert one.nit == k
        np.testing.assert_array_equal(np.array(one.allvecs), br.allvecs[i])
        np.testing.assert_array_equal(np.array(one.allfuns), br.allfuns[i])
        np.testing.assert_array_equal(np.array(one.allerrs), br.allerrs[i])
    # F and errors only; a capacity truncates every start's trace and says so
    funs = prob.minimize_proximal_gradient_batched(X0[:50], nesterov=True, tol_internal=1e-11,
                                                   return_all="funs", trace_capacity=5)
    assert funs.allvecs is None and funs.trace_truncated
    for i in range(50):
        np.testing.assert_array_equal(funs.allerrs[i], br.allerrs[i][:5])
        np.testing.assert_array_equal(funs.allfuns[i], br.allfuns[i][:6])
    with pytest.raises(ValueError):
        prob.minimize_proximal_gradient_batched(X0[:2], return_all="vecs")


def test_return_all_goes_through_the_device_in_groups(gpu, monkeypatch):
    """With a small device budget the iterates are fetched in several groups of starts: the
    result must not depend on the grouping."""
    import zfista_b200.problems as zp
    from zfista_b200 import proximal_gradient as pg

    prob = zp.FDS(n_features=40, l1_ratios=(np.arange(3) + 1) / 40, l1_shifts=np.arange(3.0))
    X0 = np.random.RandomState(4).uniform(-2, 2, size=(37, 40))
    kw = dict(nesterov=True, tol_internal=1e-11, return_all=True)
    whole = prob.minimize_proximal_gradient_batched(X0, **kw)
    monkeypatch.setattr(pg, "_TRACE_DEVICE_BYTES", 200_000)
    parts = prob.minimize_proximal_gradient_batched(X0, **kw)
    np.testing.assert_array_equal(parts.nit, whole.nit)
    np.testing.assert_array_equal(parts.allvecs.flat, whole.allvecs.flat)
    np.testing.assert_array_equal(parts.allfuns.flat, whole.allfuns.flat)
    np.testing.assert_array_equal(parts.allerrs.flat, whole.allerrs.flat)
    # the single-objective problem classes return shape-(1,) objective arrays, as the reference's do
    lfr = zp.LinearFunctionRank1(n_features=6, n_objectives=1)
    one = lfr.minimize_proximal_gradient(np.ones(6) * 0.3, return_all=True, max_iter=20)
    assert np.shape(one.fun) == (1,) and np.shape(one.allfuns[0]) == (1,)


def test_host_entry_points_from_several_threads(gpu):
    """The *_host entry points keep their device scratch and their stream per calling thread
    (round 1: one process-wide arena behind a mutex on the legacy stream): concurrent calls from
    four threads -- ctypes releases the GIL -- give exactly the serial results."""
    from concurrent.futures import ThreadPoolExecutor

    import zfista_b200.problems as zp

    probs = [zp.JOS1(n_features=30), zp.FDS(n_features=40, l1_ratios=(np.arange(3) + 1) / 40,
                                           l1_shifts=np.arange(3.0)),
             zp.JOS1(n_features=8, l1_ratios=(0.1, 0.2), l1_shifts=(0.0, 1.0)), zp.SD()]
    rng = np.random.RandomState(21)
    starts = [rng.uniform(0.5, 2.0, size=(200 + 37 * k, p.n_features)) for k, p in enumerate(probs)]
    kw = dict(nesterov=True, tol_internal=1e-11)
    serial = [p.minimize_proximal_gradient_batched(X, **kw) for p, X in zip(probs, starts)]

    def work(k):
        out = []
        for _ in range(4):
            out.append(probs[k].minimize_proximal_gradient_batched(starts[k], **kw))
        return out

    with ThreadPoolExecutor(max_workers=4) as ex:
        results = list(ex.map(work, range(4)))
    for k in range(4):
        for br in results[k]:
            np.testing.assert_array_equal(br.nit, serial[k].nit)
            np.testing.assert_array_equal(br.x, serial[k].x)
            np.testing.assert_array_equal(br.fun, serial[k].fun)
