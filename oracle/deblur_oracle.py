"""CPU oracle for the cameraman-style deblurring workload.  TEST INFRASTRUCTURE ONLY.

Restates the closures of /root/reference/examples/cameraman.ipynb (cell "Objective
function"):

    f(x)      = || R W x - b ||^2          R = correlate2d(., K, mode="same", boundary="symm")
    g(x)      = l1 * ||x||_1               W = inverse single-level 2-D Haar transform
    jac_f(x)  = 2 W^T R (R W x - b)        (the notebook applies R again for R^T)
    prox(w,x) = soft-threshold(x, l1 * w)

PARITY:
* the SOLVER on this path is pinned: tests/golden/deblur_*.npz were produced by the
  unmodified reference ``zfista.minimize_proximal_gradient`` driven by these closures
  (tests/golden/make_golden.py: gen_deblur_cases);
* R is scipy.signal.correlate2d itself (scipy is in the image);
* W / W^T are a restatement: the notebook uses ``pywt.dwt2 / idwt2(.., "haar")``
  (PyWavelets, a notebook-only dependency that is NOT installed here and not listed in
  pyproject.toml; any 1.x release).  Published definition used: dec_lo = [1, 1]/sqrt(2),
  dec_hi = [-1, 1]/sqrt(2) applied with downsampling, i.e. per 2x2 block
        cA = (p00 + p01 + p10 + p11)/2     cH = (p00 + p01 - p10 - p11)/2   (detail on axis 0)
        cV = (p00 - p01 + p10 - p11)/2     cD = (p00 - p01 - p10 + p11)/2
  coefficient vector = [cA, cH, cV, cD].flatten()  (dwt_array in the notebook).
  "parity unpinned" for the sign convention of cH / cV / cD only: ||x||_1, F and the
  iteration counts do not depend on it.
* ``skimage.filters.window(("gaussian", 4), (9, 9))`` (scikit-image, absent) is an input
  here: any odd-sized kernel K.  `gaussian_kernel` gives the radial Gaussian the notebook's
  call approximates.
"""
from __future__ import annotations

import numpy as np
from scipy.signal import correlate2d


def gaussian_kernel(size=9, sigma=4.0):
    c = (size - 1) / 2
    i = np.arange(size) - c
    return np.exp(-(i[:, None] ** 2 + i[None, :] ** 2) / (2 * sigma ** 2))


def dwt_array(image):
    p00, p01 = image[0::2, 0::2], image[0::2, 1::2]
    p10, p11 = image[1::2, 0::2], image[1::2, 1::2]
    cA = (p00 + p01 + p10 + p11) / 2
    cH = (p00 + p01 - p10 - p11) / 2
    cV = (p00 - p01 + p10 - p11) / 2
    cD = (p00 - p01 - p10 + p11) / 2
    return np.array([cA, cH, cV, cD]).flatten()


def idwt_array(array, shape):
    h, w = shape
    cA, cH, cV, cD = np.asarray(array).reshape(4, h // 2, w // 2)
    out = np.empty((h, w))
    out[0::2, 0::2] = (cA + cH + cV + cD) / 2
    out[0::2, 1::2] = (cA + cH - cV - cD) / 2
    out[1::2, 0::2] = (cA - cH + cV - cD) / 2
    out[1::2, 1::2] = (cA - cH - cV + cD) / 2
    return out


def blur(image, kernel):
    return correlate2d(image, kernel, mode="same", boundary="symm")


def lipschitz(kernel):
    """The notebook's L = 2 * max|dctn(K) / dctn(unit)|^2 (Hansen et al. 2006)."""
    from scipy.fftpack import dctn

    unit = np.zeros(kernel.shape)
    unit[0, 0] = 1
    spectrum = dctn(kernel) / dctn(unit)
    return 2 * np.max(np.abs(spectrum)) ** 2


def closures(observed, kernel, l1_ratio):
    """(f, g, jac_f, prox_wsum_g) exactly as the notebook defines them."""
    shape = observed.shape

    def f(x):
        return np.array([np.linalg.norm(blur(idwt_array(x, shape), kernel) - observed) ** 2])

    def jac_f(x):
        return 2 * dwt_array(
            blur(blur(idwt_array(x, shape), kernel) - observed, kernel)).reshape(1, -1)

    def g(x):
        return np.array([l1_ratio * np.linalg.norm(x, ord=1)])

    def prox_wsum_g(weight, x):
        return np.where(np.abs(x) <= l1_ratio * weight, 0, x - l1_ratio * weight * np.sign(x))

    return f, g, jac_f, prox_wsum_g


def synthetic_scene(h, w, seed=0, noise=1e-3, kernel=None):
    """A piecewise-smooth test image in [0, 1] (stand-in for skimage.data.camera()[::2, ::2]
    / 255), its blurred + noisy observation, and the kernel."""
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = 0.5 + 0.25 * np.sin(2 * np.pi * xx / w * 1.5) * np.cos(2 * np.pi * yy / h)
    for _ in range(6):
        cy, cx = rng.uniform(0, h), rng.uniform(0, w)
        ry, rx = rng.uniform(h / 16, h / 4), rng.uniform(w / 16, w / 4)
        img[((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 < 1] = rng.uniform(0, 1)
    img = np.clip(img, 0, 1)
    kernel = gaussian_kernel() if kernel is None else kernel
    observed = blur(img, kernel) + rng.standard_normal((h, w)) * noise
    return img, observed, kernel


def minimize(observed, kernel, l1_ratio, x0, **kwargs):
    """Oracle solve: zfista_oracle's restatement of the reference loop on these closures."""
    from . import zfista_oracle as zo

    f, g, jac_f, prox = closures(observed, kernel, l1_ratio)
    spec = zo.ProblemSpec("Closures", x0.shape[0], 1,
                          extra=dict(f=f, g=g, jac_f=jac_f, prox_wsum_g=prox))
    return zo.minimize_proximal_gradient(spec, x0, **kwargs)
