"""CPU: the deblurring oracle against the fixtures made by the unmodified reference solver
driven by the notebook's closures (tests/golden/make_golden.py: gen_deblur_cases), and the
Haar restatement's algebraic properties."""
import warnings

import numpy as np
import pytest

import helpers
from oracle import deblur_oracle as do


def test_haar_pair_is_orthonormal_and_blur_matches_definition():
    rng = np.random.RandomState(0)
    img = rng.standard_normal((12, 20))
    x = do.dwt_array(img)
    np.testing.assert_allclose(do.idwt_array(x, img.shape), img, atol=1e-15)
    np.testing.assert_allclose(np.linalg.norm(x), np.linalg.norm(img), rtol=1e-14)
    # 1-D convention: pywt.dwt([1, 2], "haar") = (2.1213.., -0.7071..)  => cV of [[1,2],[1,2]] < 0
    c = do.dwt_array(np.array([[1.0, 2.0], [1.0, 2.0]]))
    np.testing.assert_allclose(c, [3.0, 0.0, -1.0, 0.0])
    # correlate2d "symm": explicit reflected padding
    K = rng.standard_normal((5, 5))
    pad = np.pad(img, 2, mode="symmetric")
    ref = np.array([[np.sum(pad[i:i + 5, j:j + 5] * K) for j in range(20)] for i in range(12)])
    np.testing.assert_allclose(do.blur(img, K), ref, atol=1e-13)


@pytest.mark.parametrize("tag", ["s32", "s48x64"])
def test_oracle_reproduces_reference_runs(tag):
    d = helpers.load("deblur")
    obs, kernel, l1 = d[f"{tag}_observed"], d[f"{tag}_kernel"], float(d[f"{tag}_l1"])
    x0, L = d[f"{tag}_x0"], float(d[f"{tag}_L"])
    np.testing.assert_array_equal(do.dwt_array(obs), x0)
    assert do.lipschitz(kernel) == L
    f, g, jac_f, _ = do.closures(obs, kernel, l1)
    for k, x in enumerate(d[f"{tag}_evalX"]):
        assert f(x)[0] == d[f"{tag}_evalf"][k]
        np.testing.assert_array_equal(jac_f(x)[0], d[f"{tag}_evaljac"][k])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in (0, 3):
            ab = tuple(d[f"{tag}_pairs"][i])
            r = do.minimize(obs, kernel, l1, x0, lr=1 / L, decay_rate=1, nesterov=True,
                            nesterov_ratio=ab, return_all=True, max_iter=400, tol=1e-5)
            assert r["nit"] == int(d[f"{tag}_fixed{i}_nit"])
            np.testing.assert_array_equal(r["x"], d[f"{tag}_fixed{i}_x"])
            np.testing.assert_array_equal(np.ravel(r["allerrs"]), d[f"{tag}_fixed{i}_allerrs"])
        r = do.minimize(obs, kernel, l1, x0, nesterov=True, return_all=True, max_iter=150)
        assert r["nit"] == int(d[f"{tag}_bt_fista_nit"])
        np.testing.assert_array_equal(r["x"], d[f"{tag}_bt_fista_x"])
