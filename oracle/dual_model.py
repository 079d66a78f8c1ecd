"""CPU model of the two DEVICE-SIDE dual solvers.  TEST INFRASTRUCTURE ONLY.

The CUDA kernels (zfista_b200/csrc/zf_dual.cuh) cannot call scipy, so they carry
their own fp64 solvers for the dual of the proximal subproblem
(proximal_gradient.py:61-76).  This file states the same two algorithms in
Python so tests can check them, step for step, against what scipy returns on the
golden subproblems -- before and independently of any GPU run:

* :func:`fmin_bounded` -- the bounded Brent / golden-section minimiser
  ("fmin", Forsythe, Malcolm & Moler 1977; Brent 1973), with the exact
  tolerances and update order of ``scipy.optimize._optimize.
  _minimize_scalar_bounded`` (scipy 1.18.1), which is what the reference's
  ``minimize_scalar(bounds=(0, 1), options={"xatol": tol})`` call runs for two
  objectives.  The device follows this sequence so that it lands on the same
  weight, including the ~sqrt(eps) termination offsets the reference has at the
  ends of [0, 1].

* :func:`simplex_newton` -- for three or more objectives the reference calls
  trust-constr (an interior-point method; not reproducible step by step).  The
  dual is a concave piecewise-quadratic over the simplex, so the device solves
  it exactly: semi-smooth Newton steps whose QP over the m-simplex (m <= 4) is
  solved by enumerating faces in registers.  Matching is then "both converge to
  the same optimum" (trust-constr at gtol = xtol = barrier_tol = tol_internal).
"""
from __future__ import annotations

import itertools
import math

import numpy as np

SQRT_EPS = math.sqrt(2.2e-16)
GOLDEN = 0.5 * (3.0 - math.sqrt(5.0))


def _sgn(v):
    return float(v > 0) - float(v < 0)


def fmin_bounded(func, a=0.0, b=1.0, xatol=1e-12, maxfun=100000):
    """Minimise func on [a, b]; returns (x, f(x), nfev)."""
    fulc = a + GOLDEN * (b - a)
    nfc = xf = fulc
    rat = e = 0.0
    x = xf
    fx = func(x)
    num = 1
    ffulc = fnfc = fx
    xm = 0.5 * (a + b)
    tol1 = SQRT_EPS * abs(xf) + xatol / 3.0
    tol2 = 2.0 * tol1
    while abs(xf - xm) > (tol2 - 0.5 * (b - a)):
        golden = True
        if abs(e) > tol1:
            golden = False
            r = (xf - nfc) * (fx - ffulc)
            q = (xf - fulc) * (fx - fnfc)
            p = (xf - fulc) * q - (xf - nfc) * r
            q = 2.0 * (q - r)
            if q > 0.0:
                p = -p
            q = abs(q)
            r = e
            e = rat
            if (abs(p) < abs(0.5 * q * r)) and (p > q * (a - xf)) and (p < q * (b - xf)):
                rat = (p + 0.0) / q
                x = xf + rat
                if ((x - a) < tol2) or ((b - x) < tol2):
                    si = _sgn(xm - xf) + ((xm - xf) == 0)
                    rat = tol1 * si
            else:
                golden = True
        if golden:
            e = (a - xf) if xf >= xm else (b - xf)
            rat = GOLDEN * e
        si = _sgn(rat) + (rat == 0)
        x = xf + si * max(abs(rat), tol1)
        fu = func(x)
        num += 1
        if fu <= fx:
            if x >= xf:
                a = xf
            else:
                b = xf
            fulc, ffulc = nfc, fnfc
            nfc, fnfc = xf, fx
            xf, fx = x, fu
        else:
            if x < xf:
                a = x
            else:
                b = x
            if (fu <= fnfc) or (nfc == xf):
                fulc, ffulc = nfc, fnfc
                nfc, fnfc = x, fu
            elif (fu <= ffulc) or (fulc == xf) or (fulc == nfc):
                fulc, ffulc = x, fu
        xm = 0.5 * (a + b)
        tol1 = SQRT_EPS * abs(xf) + xatol / 3.0
        tol2 = 2.0 * tol1
        if num >= maxfun:
            break
    return xf, fx, num


# --------------------------------------------------------------------------
# dual pieces
# --------------------------------------------------------------------------
def prox_chain(v, coef, shifts, lower, upper, has_l1):
    """The reference's prox chain (problems.py:126-137) on the vector v, returning
    also, per coordinate, alpha (1 if p moves with v, 0 if pinned at a kink or a
    bound), eps[i] = +1/-1: side of shift i the free coordinate lies on, and the piece code
    of zf_problems.cuh:prox_elem (two bits per L1 stage: 2 pinned at its kink, 1 / 0 above /
    below it; two bits for the box: 1 clipped at the upper bound, 2 at the lower)."""
    n = v.shape[0]
    m = coef.shape[0]
    alpha = np.ones(n)
    eps = np.zeros((m, n))
    code = np.zeros(n, dtype=np.int64)
    p = v
    if has_l1:
        a0 = p + np.sum(coef[1:]) - shifts[0] + shifts[0]
        p = np.sign(a0) * np.maximum(np.abs(a0) - coef[0], 0.0)
        stuck = np.abs(a0) <= coef[0]
        alpha = np.where(stuck, 0.0, alpha)
        eps[0] = np.where(a0 > coef[0], 1.0, -1.0)
        code |= np.where(stuck, 2, np.where(a0 > coef[0], 1, 0))
        for i in range(1, m):
            ai = p - coef[i] - shifts[i]
            p = np.sign(ai) * np.maximum(np.abs(ai) - coef[i], 0.0) + shifts[i]
            stuck = np.abs(ai) <= coef[i]
            alpha = np.where(stuck, 0.0, alpha)
            eps[i] = np.where(ai > coef[i], 1.0, -1.0)
            code |= np.where(stuck, 2, np.where(ai > coef[i], 1, 0)) << (2 * i)
    if lower is not None:
        q = np.clip(p, lower, upper)
        alpha = np.where(q != p, 0.0, alpha)
        code |= np.where(q < p, 1, np.where(q > p, 2, 0)) << (2 * m)
        p = q
    return p, alpha, eps, code


def dual_eval(w, y, J, lr, c, l1_ratios, l1_shifts, lower, upper, g_fun):
    """D(w), grad D(w), Q = lr * M M^T  (minus the generalised Hessian)."""
    has_l1 = l1_ratios is not None
    m = w.shape[0]
    wj = w @ J
    v = y - lr * wj
    lam = l1_ratios if has_l1 else np.zeros(m)
    sh = l1_shifts if has_l1 else np.zeros(m)
    p, alpha, eps, _ = prox_chain(v, lr * w * lam, sh, lower, upper, has_l1)
    gp = g_fun(p)
    D = np.inner(w, gp) + np.sum((p - v) ** 2) / 2 / lr - lr / 2 * np.sum(wj ** 2) \
        + np.inner(w, c)
    G = gp + J @ (p - y) + c
    M = (J + lam[:, None] * eps) * alpha
    Q = lr * (M @ M.T)
    return D, G, Q, p


def simplex_qp(Q, G, w_cur):
    """argmax over the unit simplex of the local model  G.d - 0.5 d'Qd,  d = w' - w_cur
    (Q PSD, m <= 4) by face enumeration; singular faces are skipped (their optimum
    lies on a sub-face).  Everything is expressed in the step d: the gradient G is
    O(1) while Q w can be ~1e12 (FDS, n = 100), so forming G + Q w first would wipe
    out the low bits of G that decide the optimum.

    Faces are visited in the device's order (zf_dual.cuh:QpFaces): support masks from the
    whole simplex (2^m - 1) down to 1, and if the maximiser over the whole simplex's affine
    hull is feasible it is the global maximiser and the sub-faces are not visited."""
    m = G.shape[0]
    best_w, best_val = w_cur.copy(), -np.inf
    scale = np.trace(Q)
    for mask in range((1 << m) - 1, 0, -1):
        S = [b for b in range(m) if mask & (1 << b)]
        k = len(S)
        w = np.zeros(m)
        e0 = np.zeros(m)
        s0 = S[0]
        e0[s0] = 1.0
        if k == 1:
            w = e0
        else:
            u = Q @ (e0 - w_cur)
            R = np.empty((k - 1, k - 1))
            rhs = np.empty(k - 1)
            for a in range(k - 1):
                ia = S[a + 1]
                rhs[a] = (G[ia] - u[ia]) - (G[s0] - u[s0])
                for b in range(k - 1):
                    ib = S[b + 1]
                    R[a, b] = Q[ia, ib] - Q[ia, s0] - Q[s0, ib] + Q[s0, s0]
            # LDL^T without pivoting; a pivot at rounding level = singular face
            A = R.copy()
            ok = True
            inv = np.zeros(k - 1)
            for kk in range(k - 1):
                if not (A[kk, kk] > 1e-13 * scale):
                    ok = False
                    break
                inv[kk] = 1.0 / A[kk, kk]
                for r in range(kk + 1, k - 1):
                    fct = A[r, kk] * inv[kk]
                    A[r, kk + 1:] -= fct * A[kk, kk + 1:]
                    rhs[r] -= fct * rhs[kk]
            if not ok:
                continue
            z = np.zeros(k - 1)
            for kk in range(k - 2, -1, -1):
                z[kk] = (rhs[kk] - A[kk, kk + 1:] @ z[kk + 1:]) * inv[kk]   # as zf_dual.cuh:qp_face
            if (z < 0).any() or 1.0 - z.sum() < 0:
                continue
            w[S[1:]] = z
            w[s0] = 1.0 - z.sum()
        d = w - w_cur
        val = sum(d[i] * (G[i] - 0.5 * (Q[i] @ d)) for i in range(m))
        if val > best_val:
            best_val, best_w = val, w
        if mask == (1 << m) - 1 and best_val > -np.inf:
            break
    return best_w


def simplex_newton(y, J, lr, c, l1_ratios, l1_shifts, lower, upper, g_fun, w0=None,
                   max_iter=60, shortcut=True, info=None):
    """Maximise the dual over the simplex.  Returns (w, D(w), p(w), dual evaluations).

    Statement, step for step, of zf_dual.cuh:dual_newton.  Each step solves the QP of the
    current quadratic piece exactly.  When the model-predicted gain of a step is below the
    rounding level of D the step is taken on trust and the iteration stops (a Newton step is
    accurate far below what a comparison of D values can resolve).  ``shortcut``: the dual is
    ONE quadratic while no coordinate changes its piece of the prox chain, so if the primal
    point of the Newton candidate lies on the same pieces as the point just evaluated the
    candidate IS the maximiser and D(candidate) = D + pred; the confirming evaluation is
    skipped and the primal point of the probe is returned."""
    m = J.shape[0]
    has_l1 = l1_ratios is not None
    lam = l1_ratios if has_l1 else np.zeros(m)
    sh = l1_shifts if has_l1 else np.zeros(m)
    w = np.ones(m) / m if w0 is None else np.array(w0, dtype=np.float64)
    args = (y, J, lr, c, l1_ratios, l1_shifts, lower, upper, g_fun)

    def full(wq):
        Dq, Gq, Qq, pq = dual_eval(wq, *args)
        vq = y - lr * (wq @ J)
        _, _, _, codes = prox_chain(vq, lr * wq * lam, sh, lower, upper, has_l1)
        return Dq, Gq, Qq, codes

    wt = w.copy()
    d = np.zeros(m)
    step, evals, it, bt, first = 1.0, 0, 0, 0, True
    D = G = Q = pat = None
    x_ready = None
    while True:
        Dt, Gt, Qt, patt = full(wt)
        evals += 1
        if first or Dt >= D:
            first = False
            w, D, G, Q, pat = wt.copy(), Dt, Gt, Qt, patt
        else:
            # the device overwrites the stored piece codes at every full evaluation, also at a
            # rejected trial point
            pat = patt
            step *= 0.5
            bt += 1
            if bt >= 30:
                break
            wt = w + step * d
            continue
        it += 1
        if it > max_iter:
            break
        wn = simplex_qp(Q, G, w)
        d = wn - w
        if np.max(np.abs(d)) == 0.0:
            break
        pred = G @ d - 0.5 * d @ Q @ d
        if pred <= 1e-15 * (abs(D) + np.max(np.abs(G))):
            w = wn
            D = D + pred
            break
        if shortcut:
            vn = y - lr * (wn @ J)
            pn, _, _, codes = prox_chain(vn, lr * wn * lam, sh, lower, upper, has_l1)
            if np.array_equal(codes, pat):
                w = wn
                D = D + pred
                x_ready = pn
                break
        step, bt = 1.0, 0
        wt = wn.copy()
    if x_ready is None:
        v = y - lr * (w @ J)
        x_ready, _, _, _ = prox_chain(v, lr * w * lam, sh, lower, upper, has_l1)
    if info is not None:
        info["evals"] = evals
    return w, D, x_ready, evals
