"""CPU tests: pin the oracle to the real reference.

Fixtures under tests/golden/ were produced by running the unmodified reference
(tests/golden/make_golden.py).  These tests replay the same inputs through
oracle/zfista_oracle.py and require equality, so that the oracle can stand in for
the reference on the GPU box (where /root/reference does not exist).
"""
import warnings

import numpy as np
import pytest
from scipy.optimize import minimize_scalar

import helpers
from oracle import dual_model as dm
from oracle import zfista_oracle as zo

FIX = helpers.fixture_problems()


# ---------------------------------------------------------------- problem classes
@pytest.mark.parametrize("pname", sorted(FIX))
def test_problem_functions_match_reference(pname):
    d = helpers.load("problem_eval")
    cls, kw = FIX[pname]
    spec = helpers.oracle_spec(cls, kw)
    X, W = d[pname + "__X"], d[pname + "__W"]
    for k in range(len(X)):
        np.testing.assert_array_equal(zo.f(spec, X[k]), d[pname + "__f"][k])
        np.testing.assert_array_equal(zo.g(spec, X[k]), d[pname + "__g"][k])
        np.testing.assert_array_equal(zo.jac_f(spec, X[k]), d[pname + "__jac"][k])
        np.testing.assert_array_equal(zo.prox_wsum_g(spec, W[k], X[k]), d[pname + "__prox"][k])


def test_reference_unit_vectors():
    """Known answers of the reference's own tests/test_problems.py."""
    x = np.array([1, 2, 3, 4, 5])
    jos1 = zo.make_spec("JOS1")
    np.testing.assert_almost_equal(zo.f(jos1, x), [11, 3])
    np.testing.assert_almost_equal(
        zo.jac_f(jos1, x), [[2 / 5, 4 / 5, 6 / 5, 8 / 5, 2], [-2 / 5, 0, 2 / 5, 4 / 5, 6 / 5]])
    jl1 = zo.make_spec("JOS1", l1_ratios=[0.2, 0.1], l1_shifts=[0, 1])
    np.testing.assert_almost_equal(zo.g(jl1, x), [3, 1])
    np.testing.assert_almost_equal(
        zo.prox_wsum_g(jl1, np.array([0.5, 0.5]), np.array([3, 4, 5, 6, 7])),
        [2.85, 3.85, 4.85, 5.85, 6.85])
    sd = zo.make_spec("SD")
    xs = np.array([1, np.sqrt(2), np.sqrt(2), 1])
    np.testing.assert_almost_equal(zo.f(sd, xs), [7, 8])
    np.testing.assert_almost_equal(
        zo.jac_f(sd, xs), [[2, np.sqrt(2), np.sqrt(2), 1], [-2, -np.sqrt(2), -np.sqrt(2), -2]])
    np.testing.assert_almost_equal(zo.g(sd, xs), [0, 0])
    np.testing.assert_almost_equal(zo.prox_wsum_g(sd, np.array([0.5, 0.5]), xs), xs)
    fds = zo.make_spec("FDS", n_features=5)
    np.testing.assert_almost_equal(zo.f(fds, x), [0.0, 75.0855369, 0.1183459])
    np.testing.assert_almost_equal(
        zo.jac_f(fds, x),
        [[0, 0, 0, 0, 0],
         [6.01710738, 8.01710738, 10.0171074, 12.0171074, 14.0171074],
         [-0.0613132402, -0.0360894089, -0.0149361205, -4.88417037e-03, -1.12299117e-03]])
    fc = zo.make_spec("FDS", n_features=5, bounds=(0, np.inf))
    np.testing.assert_almost_equal(zo.g(fc, np.ones(5)), [0, 0, 0])
    assert np.all(np.isinf(zo.g(fc, -np.ones(5))))
    np.testing.assert_almost_equal(
        zo.prox_wsum_g(fc, np.ones(3) / 3, np.array([-3, -1, 0, 1, 3])), [0, 0, 0, 1, 3])


# ---------------------------------------------------------------- inner solvers
def test_bounded_brent_is_scipy_bit_for_bit():
    rng = np.random.RandomState(0)
    for t in range(150):
        a, b, c = rng.uniform(0.1, 3), rng.uniform(-3, 3), rng.uniform(-1, 1)
        fn = [lambda w: a * (w - b) ** 2 + c,
              lambda w: a * abs(w - b / 3) + 0.3 * (w - c) ** 2,
              lambda w: np.exp(a * w) - b * w][t % 3]
        r = minimize_scalar(fn, bounds=(0, 1), options={"maxiter": 100000, "xatol": 1e-11})
        xf, fx, n = dm.fmin_bounded(fn, 0.0, 1.0, xatol=1e-11, maxfun=100000)
        assert (xf, fx, n) == (r.x, r.fun, r.nfev)


@pytest.mark.parametrize("pname", ["JOS1_n5", "JOS1_n5_l1", "JOS1_n50_l1", "SD", "ZDT1_n50",
                                   "TOI4_l1", "TRIDIA", "TRIDIA_l1", "LFR1_n30", "FDS_n10",
                                   "FDS_n10_l1", "FDS_n10_box"])
def test_subproblem_matches_reference(pname):
    """oracle.solve_subproblem == reference _solve_subproblem on the stored inputs."""
    d = helpers.load("subproblem")
    cls, kw = FIX[pname]
    spec = helpers.oracle_spec(cls, kw)
    m = spec.n_objectives
    Y, XO, LR = d[pname + "__y"], d[pname + "__xold"], d[pname + "__lr"]
    ks = range(len(Y)) if m == 2 else range(0, len(Y), 4)   # trust-constr is slow
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for k in ks:
            x, fun, w, _ = zo.solve_subproblem(spec, LR[k], XO[k], Y[k], np.ones(m) / m,
                                               tol=1e-11, max_iter=100000,
                                               deprecated=(k % 4 == 3))
            np.testing.assert_array_equal(x, d[pname + "__x"][k])
            assert fun == d[pname + "__fun"][k]
            np.testing.assert_array_equal(w, d[pname + "__w"][k])


def _dual_args(spec, y, xold, lr, deprecated):
    fy = zo.f(spec, y)
    Fp = zo.f(spec, xold) + zo.g(spec, xold)
    J = zo.jac_f(spec, y)
    c = np.zeros_like(fy) if deprecated else fy - Fp
    return (y, J, lr, c, spec.l1_ratios, spec.l1_shifts, spec.lower, spec.upper,
            lambda p: zo.g(spec, p))


@pytest.mark.parametrize("pname", ["JOS1_n5_l1", "SD", "ZDT1_n50", "TOI4_l1", "TRIDIA",
                                   "TRIDIA_l1", "LFR1_n30", "FDS_n10", "FDS_n10_l1",
                                   "FDS_n10_box", "FDS_n100_l1"])
def test_simplex_newton_never_worse_than_reference(pname):
    """The device's exact dual solver reaches a dual value >= the one scipy reached on
    every stored subproblem, with a vanishing optimality gap; for two objectives the
    weights agree with bounded Brent to its own sqrt(eps) resolution."""
    d = helpers.load("subproblem")
    cls, kw = FIX[pname]
    spec = helpers.oracle_spec(cls, kw)
    Y, XO, LR, W = (d[pname + "__" + k] for k in ("y", "xold", "lr", "w"))
    for k in range(len(Y)):
        args = _dual_args(spec, Y[k], XO[k], LR[k], k % 4 == 3)
        w, D, p, its = dm.simplex_newton(*args)
        D_ref = dm.dual_eval(W[k], *args)[0]
        Dn, G, Q, _ = dm.dual_eval(w, *args)
        scale = abs(Dn) + np.max(np.abs(G))
        assert Dn >= D_ref - 1e-13 * scale
        # optimality: one more exact QP step on the local model cannot gain anything
        # (the Frank-Wolfe gap itself is not scale free: Q reaches 1e12 for FDS n=100)
        step = dm.simplex_qp(Q, G, w) - w
        assert G @ step - 0.5 * step @ Q @ step <= 1e-13 * scale
        assert its <= 12
        assert abs(w.sum() - 1) < 1e-14 and (w >= 0).all()
        if spec.n_objectives == 2:
            np.testing.assert_allclose(w, W[k], atol=5e-8)
            np.testing.assert_allclose(p, d[pname + "__x"][k], atol=2e-7)


# ---------------------------------------------------------------- full solves
def _cheap_cases():
    out = []
    for name in helpers.golden_cases():
        pname = name.split("__")[0]
        if pname in FIX and FIX[pname][0] in ("JOS1", "SD", "ZDT1", "TOI4"):
            out.append(name)
    return out


@pytest.mark.parametrize("case", _cheap_cases())
def test_two_objective_solves_match_reference(case):
    """Full solves, two objectives: the oracle (scipy route) and the device-model
    route (restated Brent) both reproduce the reference: same nit, same x."""
    d = helpers.load(case)
    spec = helpers.oracle_spec(str(d["problem"]), helpers.case_kwargs(d))
    opts = helpers.case_options(d)
    n_check = min(len(d["x0"]), 3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(n_check):
            r = zo.minimize_proximal_gradient(spec, d["x0"][i], return_all=(i == 0), **opts)
            assert r["nit"] == d["nit"][i]
            assert r["success"] == d["success"][i]
            np.testing.assert_array_equal(r["x"], d["x"][i])
            np.testing.assert_array_equal(r["fun"], d["fun"][i])
            if i == 0:
                np.testing.assert_array_equal(np.array(r["allerrs"]), d["allerrs0"])
                np.testing.assert_array_equal(np.array(r["allfuns"]), d["allfuns0"])
            r2 = zo.minimize_proximal_gradient(
                spec, d["x0"][i], subproblem=zo.solve_subproblem_device_model, **opts)
            assert r2["nit"] == d["nit"][i]
            np.testing.assert_allclose(r2["x"], d["x"][i], rtol=0, atol=1e-12)


def test_three_objective_solve_matches_reference():
    """One cheap tri-objective case through the scipy (trust-constr) route."""
    d = helpers.load("TRIDIA__ista")
    spec = helpers.oracle_spec("TRIDIA", {})
    opts = helpers.case_options(d)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = zo.minimize_proximal_gradient(spec, d["x0"][0], **opts)
    assert r["nit"] == d["nit"][0]
    np.testing.assert_array_equal(r["x"], d["x"][0])


def test_lasso_single_objective_matches_reference():
    d = helpers.load("lasso_single")
    # toy problems of tests/test_proximal_gradient.py:66-111 and their asserted optima
    expected = {1e-8: 1.0, 0.1: 0.85, 0.5: 0.25, 1.0: 0.0}
    for l1, nest, x_ref, fun_ref, nit_ref in d["toy_rows"]:
        spec = zo.make_least_squares_l1(d["toy_A"], d["toy_b"], l1, scale=1 / 6)
        r = zo.minimize_proximal_gradient(spec, d["toy_x0"], nesterov=bool(nest))
        assert r["nit"] == int(nit_ref)
        assert r["x"][0] == x_ref
        np.testing.assert_almost_equal(r["x"], [expected[float(l1)]], decimal=3)
    spec = zo.make_least_squares_l1(d["ds_A"], d["ds_b"], float(d["ds_l1"]),
                                    scale=float(d["ds_scale"]))
    L = float(d["ds_L"])
    runs = {
        "bt_ista": dict(nesterov=False),
        "bt_fista": dict(nesterov=True),
        "fixed_fista": dict(nesterov=True, lr=1 / L, decay_rate=1),
        "fixed_fista_ab": dict(nesterov=True, lr=1 / L, decay_rate=1,
                               nesterov_ratio=(0.5, 1 / 16)),
    }
    for tag, opts in runs.items():
        r = zo.minimize_proximal_gradient(spec, d["ds_x0"], return_all=True, max_iter=20000,
                                          **opts)
        assert r["nit"] == int(d[f"ds_{tag}_nit"])
        np.testing.assert_array_equal(r["x"], d[f"ds_{tag}_x"])
        np.testing.assert_array_equal(np.array(r["allerrs"]), d[f"ds_{tag}_allerrs"])


def test_lasso_ab_sweep_matches_reference():
    """The (a, b) momentum sweep over one shared A (fixture made by the unmodified reference,
    one solve per pair): the oracle reproduces every run bit for bit -- the parity anchor of
    the lockstep multi-run device path (tests/test_gpu_lasso_multi.py)."""
    d = helpers.load("lasso_ab_sweep")
    spec = zo.make_least_squares_l1(d["A"], d["b"], float(d["l1"]), scale=float(d["scale"]))
    L = float(d["L"])
    for tag, opts in {"bt": dict(), "fixed": dict(lr=1 / L, decay_rate=1)}.items():
        for k, ab in enumerate(d["grid"]):
            r = zo.minimize_proximal_gradient(spec, d["X0"][k], nesterov=True,
                                              nesterov_ratio=(float(ab[0]), float(ab[1])),
                                              max_iter=20000, **opts)
            assert r["nit"] == int(d[f"{tag}_nit"][k]), (tag, k)
            np.testing.assert_array_equal(r["x"], d[f"{tag}_x"][k])
            assert float(r["fun"]) == float(d[f"{tag}_fun"][k])
    assert np.allclose(d["grid"], helpers.AB_GRID)


def test_reference_is_rounding_sensitive_on_l1_cases():
    """Evidence for the parity tiers of DESIGN.md: with zero noise the restated Brent route
    reproduces the reference bit for bit; with ONE ulp of relative noise on the dual value
    the reference's own iteration counts and final points move by far more than 1e-8 on an
    L1 case, and not at all on a smooth one."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        d = helpers.load("JOS1_n50_l1__fista")
        spec = helpers.oracle_spec("JOS1", helpers.case_kwargs(d))
        opts = helpers.case_options(d)
        r0 = zo.minimize_proximal_gradient(
            spec, d["x0"][0], subproblem=helpers.noisy_brent_subproblem(0.0, 0), **opts)
        assert r0["nit"] == d["nit"][0]
        np.testing.assert_allclose(r0["x"], d["x"][0], rtol=0, atol=1e-12)
        env = helpers.oracle_noise_envelope(spec, d["x0"], d["x"], d["fun"], d["nit"], opts,
                                            n_starts=3, seeds=(0, 1))
        assert env["dnit"] >= 5 and env["dx"] > 1e-5
        d = helpers.load("JOS1_n50__fista")
        spec = helpers.oracle_spec("JOS1", helpers.case_kwargs(d))
        env = helpers.oracle_noise_envelope(spec, d["x0"], d["x"], d["fun"], d["nit"],
                                            helpers.case_options(d), n_starts=2, seeds=(0,))
        assert env["dnit"] == 0 and env["dx"] < 1e-8


# ---------------------------------------------------------------- m >= 3, converged reference
@pytest.mark.parametrize("case", helpers.converged_cases())
def test_device_model_vs_converged_reference(case):
    """The CPU statement of the device algorithm (exact simplex-Newton dual) against CONVERGED
    solves of the unmodified reference (trust-constr, default max_iter_internal;
    tests/golden/make_golden_converged.py).  trust-constr returns the dual weights to ~1e-6, so
    the two follow slightly different paths along the Pareto set: final F within 3.6e-3 relative
    on every start, the first 10 iterations of F within 1.7e-2, iteration counts within a factor
    2.5 (identical on 25 of the 78 stored starts).  tests/test_gpu_multiobjective.py makes the
    same comparison with the CUDA kernel in the model's place."""
    d = helpers.load_converged(case)
    cls, kw, opts = str(d["problem"]), helpers.case_kwargs(d), helpers.case_options(d)
    spec = helpers.oracle_spec(cls, kw)
    for i in np.flatnonzero(d["success"])[:3]:
        r = helpers.device_model_solve(spec, d["x0"][i], dict(opts, return_all=True))
        assert r["status"] == 1
        _, F_ref = helpers.converged_trace(d, i)
        scale = np.maximum(1.0, np.abs(d["fun"][i]))
        assert np.max(np.abs(r["fun"] - d["fun"][i]) / scale) < 5e-3
        k = min(10, int(d["nit"][i]), r["nit"])
        Fm = np.array(r["allfuns"])
        head = np.abs(Fm[:k + 1] - F_ref[:k + 1]) / np.maximum(1.0, np.abs(F_ref[:k + 1]))
        assert head.max() < 2e-2
        assert 1 / 3.5 <= r["nit"] / max(1, int(d["nit"][i])) <= 3.5
