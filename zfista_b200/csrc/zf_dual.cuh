// Device solvers for the dual of the multi-objective proximal subproblem
// (proximal_gradient.py:61-76, 161-209).  One warp solves one subproblem; the
// n-dimensional work is strided over the lanes and every scalar decision is taken
// redundantly (and identically) by all 32 lanes.
//
//   max_{w in simplex}  D(w) = sum_i w_i g_i(p) + ||p - v||^2 / (2 lr)
//                              - lr/2 ||J^T w||^2 + w . c
//   v = y - lr J^T w,   p = prox_{lr sum_i w_i g_i}(v),   c_i = f_i(y) - F_i(x^{k-1})
//
// The reference minimises -D:
//   m = 2 : scipy minimize_scalar(bounds=(0,1))      -> DualBrent   (same sequence)
//   m >= 3: scipy trust-constr on the simplex        -> DualNewton  (exact optimum)
#pragma once
#include "zf_problems.cuh"

#ifndef ZF_QP_BRANCHFREE
#define ZF_QP_BRANCHFREE 0
#endif

namespace zf {

template <int M>
struct DualData {
  double lr;
  double c[M];        // F_prev - f_y added to -D  (0 when deprecated): c[i] = f_y - F_prev
  bool use_c;         // !deprecated
};

// what the dual sweeps read of one coordinate: the Jacobian column and y_j
template <int M>
struct CoordIn {
  double J[M];
  double y;
};
template <int M>
__device__ __forceinline__ CoordIn<M> load_coord(const WarpCtx& c, int jc) {
  CoordIn<M> in;
#pragma unroll
  for (int i = 0; i < M; ++i) in.J[i] = c.J[i * c.n + jc];
  in.y = c.y[jc];
  return in;
}

// ------------------------------------------------------------------------------------
// -D(w) exactly as _dual_minimized_fun_jac (proximal_gradient.py:161-177) forms it.
// ------------------------------------------------------------------------------------
template <int KIND, int M, int GF>
__device__ double neg_dual_value(const zf_problem& P, const WarpCtx& c, const DualData<M>& d,
                                 const double (&w)[M]) {
  double wt[M];
#pragma unroll
  for (int i = 0; i < M; ++i) wt[i] = d.lr * w[i];
  constexpr bool L1 = (GF & ZF_G_L1) != 0;
  constexpr int NS = M + 2;
  double s[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) s[k] = 0.0;
  constexpr bool lsq = (KIND == ZF_LSQ_L1);
  sweep(c.n, c.lane, [&](int j, bool live) {
    const int jc = live ? j : 0;
    double wj = 0.0;
#pragma unroll
    for (int i = 0; i < M; ++i) wj += w[i] * c.J[i * c.n + jc];
    const double v = c.y[jc] - d.lr * wj;
    double alpha, eps[M];
            unsigned pcode;
    const double p = prox_elem<KIND, M, GF, false>(P, jc, v, wt, alpha, eps, pcode);
    if (lsq) {
      s[0] += msk(live, fabs(p));
    } else if (L1) {
#pragma unroll
      for (int i = 0; i < M; ++i) s[i] += msk(live, fabs(p - P.l1_shifts[i]));
    }
    s[M] += msk(live, p - v) * (p - v);
    s[M + 1] += msk(live, wj) * wj;
  });
  warp_sum_k<NS>(s);
  double wg = 0.0;
  if (lsq) {
    const double gl = P.l1 * s[0];
#pragma unroll
    for (int i = 0; i < M; ++i) wg += w[i] * gl;
  } else if (L1) {
#pragma unroll
    for (int i = 0; i < M; ++i) wg += w[i] * (P.l1_ratios[i] * s[i]);
  }
  double fun = -wg - norm_sq_like_numpy(s[M]) / 2.0 / d.lr +
               d.lr / 2.0 * norm_sq_like_numpy(s[M + 1]);
  if (d.use_c) {
    double wc = 0.0;
#pragma unroll
    for (int i = 0; i < M; ++i) wc += w[i] * (-d.c[i]);   // inner(w, F_prev - f_y)
    fun += wc;
  }
  return fun;
}

// x = prox_wsum_g(lr * w, y - lr * w @ J)   (proximal_gradient.py:206)
template <int KIND, int M, int GF>
__device__ void primal_from_weights(const zf_problem& P, const WarpCtx& c, double lr,
                                    const double (&w)[M], double* out) {
  double wt[M];
#pragma unroll
  for (int i = 0; i < M; ++i) wt[i] = lr * w[i];
  sweep3<CoordIn<M>, double>(
      c.n, c.lane, [&](int j, bool live) { return load_coord<M>(c, live ? j : 0); },
      [&](const CoordIn<M>& in, int j, bool live) {
        double wj = 0.0;
#pragma unroll
        for (int i = 0; i < M; ++i) wj += w[i] * in.J[i];
        const double v = in.y - lr * wj;
        double alpha, eps[M];
            unsigned pcode;
        return prox_elem<KIND, M, GF, false>(P, live ? j : 0, v, wt, alpha, eps, pcode);
      },
      [&](int j, bool live, double p) {
        if (live) out[j] = p;
      });
  __syncwarp();
}

// ------------------------------------------------------------------------------------
// Bounded Brent on [0, 1] ("fmin", Forsythe-Malcolm-Moler / Brent 1973) with the
// tolerances and update order of scipy's _minimize_scalar_bounded, which is what
// minimize_scalar(bounds=(0, 1), options={"maxiter", "xatol"}) runs
// (proximal_gradient.py:184-188).  Returns xf; *fmin = -D(xf); *nfev evaluations.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ double sgn_plus(double v) {
  // np.sign(v) + (v == 0)
  return (v > 0.0) ? 1.0 : ((v < 0.0) ? -1.0 : (v == 0.0 ? 1.0 : v));
}

template <int KIND, int GF>
__device__ double dual_brent(const zf_problem& P, const WarpCtx& c, const DualData<2>& d,
                             double xatol, int maxfun, double* fmin, int* nfev) {
  const double sqrt_eps = sqrt(2.2e-16);
  const double golden_mean = 0.5 * (3.0 - sqrt(5.0));
  double a = 0.0, b = 1.0;
  double fulc = a + golden_mean * (b - a);
  double nfc = fulc, xf = fulc;
  double rat = 0.0, e = 0.0;
  double x = xf;
  double w2[2] = {x, 1.0 - x};
  double fx = neg_dual_value<KIND, 2, GF>(P, c, d, w2);
  int num = 1;
  double fu = CUDART_INF;
  double ffulc = fx, fnfc = fx;
  double xm = 0.5 * (a + b);
  double tol1 = sqrt_eps * fabs(xf) + xatol / 3.0;
  double tol2 = 2.0 * tol1;
  while (fabs(xf - xm) > (tol2 - 0.5 * (b - a))) {
    bool golden = true;
    if (fabs(e) > tol1) {
      golden = false;
      double r = (xf - nfc) * (fx - ffulc);
      double q = (xf - fulc) * (fx - fnfc);
      double p = (xf - fulc) * q - (xf - nfc) * r;
      q = 2.0 * (q - r);
      if (q > 0.0) p = -p;
      q = fabs(q);
      r = e;
      e = rat;
      if ((fabs(p) < fabs(0.5 * q * r)) && (p > q * (a - xf)) && (p < q * (b - xf))) {
        rat = (p + 0.0) / q;
        x = xf + rat;
        if (((x - a) < tol2) || ((b - x) < tol2)) {
          rat = tol1 * sgn_plus(xm - xf);
        }
      } else {
        golden = true;
      }
    }
    if (golden) {
      e = (xf >= xm) ? (a - xf) : (b - xf);
      rat = golden_mean * e;
    }
    x = xf + sgn_plus(rat) * fmax(fabs(rat), tol1);
    w2[0] = x;
    w2[1] = 1.0 - x;
    fu = neg_dual_value<KIND, 2, GF>(P, c, d, w2);
    num += 1;
    if (fu <= fx) {
      if (x >= xf) a = xf; else b = xf;
      fulc = nfc; ffulc = fnfc;
      nfc = xf; fnfc = fx;
      xf = x; fx = fu;
    } else {
      if (x < xf) a = x; else b = x;
      if ((fu <= fnfc) || (nfc == xf)) {
        fulc = nfc; ffulc = fnfc;
        nfc = x; fnfc = fu;
      } else if ((fu <= ffulc) || (fulc == xf) || (fulc == nfc)) {
        fulc = x; ffulc = fu;
      }
    }
    xm = 0.5 * (a + b);
    tol1 = sqrt_eps * fabs(xf) + xatol / 3.0;
    tol2 = 2.0 * tol1;
    if (num >= maxfun) break;
  }
  *fmin = fx;
  *nfev = num;
  return xf;
}

// ------------------------------------------------------------------------------------
// Simplex semi-smooth Newton.  D is concave and piecewise quadratic; on the current
// piece  grad D = G,  hess D = -Qm  with  Qm = lr * Mm Mm^T,
//   Mm[i][j] = alpha_j (J[i][j] + l1_ratios_i eps_ij).
// One pass over the coordinates gives D, G (M sums) and Qm (M(M+1)/2 sums); the QP
//   max_{w' in simplex}  (G + Qm w) . w' - 1/2 w'^T Qm w'
// is then solved exactly in registers by enumerating the 2^M - 1 faces.
// ------------------------------------------------------------------------------------
// Which linear piece of the prox chain a coordinate is on: the code prox_elem<.., TRACK> builds
// (zf_problems.cuh): per L1 stage pinned / above / below its kink, and the box state.  The dual
// is one quadratic wherever these codes do not change.
// x = prox_wsum_g(lr * w, y - lr * w @ J) into `out`, and whether every coordinate is on the
// same piece as at the last full dual evaluation (c.pat)
template <int KIND, int M, int GF>
__device__ bool primal_probe(const zf_problem& P, const WarpCtx& c, double lr,
                             const double (&w)[M], double* out) {
  double wt[M];
#pragma unroll
  for (int i = 0; i < M; ++i) wt[i] = lr * w[i];
  int same = 1;
  struct In { CoordIn<M> c; unsigned short pat; };
  sweep3<In, double>(
      c.n, c.lane,
      [&](int j, bool live) {
        const int jc = live ? j : 0;
        return In{load_coord<M>(c, jc), c.pat[jc]};
      },
      [&](const In& in, int j, bool live) {
        double wj = 0.0;
#pragma unroll
        for (int i = 0; i < M; ++i) wj += w[i] * in.c.J[i];
        const double v = in.c.y - lr * wj;
        double alpha, eps[M];
            unsigned pcode;
        const double p = prox_elem<KIND, M, GF, true>(P, live ? j : 0, v, wt, alpha, eps, pcode);
        same &= live ? (int)((unsigned short)pcode == in.pat) : 1;
        return p;
      },
      [&](int j, bool live, double p) {
        if (live) out[j] = p;
      });
  __syncwarp();
  return __all_sync(ZF_FULL_MASK, same) != 0;
}

template <int M>
struct DualPoint {
  double D;
  double G[M];
  double Q[M][M];
};

template <int KIND, int M, int GF>
__device__ void dual_full(const zf_problem& P, const WarpCtx& c, const DualData<M>& d,
                          const double (&w)[M], DualPoint<M>& out) {
  double wt[M];
#pragma unroll
  for (int i = 0; i < M; ++i) wt[i] = d.lr * w[i];
  constexpr bool L1 = (GF & ZF_G_L1) != 0;
  constexpr int NQ = M * (M + 1) / 2;
  constexpr int NS = 2 * M + 2 + NQ;   // |p - s_i| sums, J_i.(p-y), ||p-v||^2, ||wj||^2, Q
  double s[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) s[k] = 0.0;
  constexpr bool lsq = (KIND == ZF_LSQ_L1);
  double lam[M], shift[M];
#pragma unroll
  for (int i = 0; i < M; ++i) {
    lam[i] = 0.0;
    shift[i] = 0.0;
    if (lsq) lam[i] = P.l1;
    else if (L1) { lam[i] = P.l1_ratios[i]; shift[i] = P.l1_shifts[i]; }
  }
  sweep3<CoordIn<M>, unsigned short>(
      c.n, c.lane, [&](int j, bool live) { return load_coord<M>(c, live ? j : 0); },
      [&](const CoordIn<M>& in, int j, bool live) {
        double wj = 0.0;
#pragma unroll
        for (int i = 0; i < M; ++i) wj += w[i] * in.J[i];
        const double yj = in.y;
        const double v = yj - d.lr * wj;
        double alpha, eps[M];
            unsigned pcode;
        const double p = prox_elem<KIND, M, GF, true>(P, live ? j : 0, v, wt, alpha, eps, pcode);
        // a dead slot contributes exact zeros: its alpha, p - y, p - v and w.J are masked
        const double am = msk(live, alpha);
        const double dy = msk(live, p - yj);
        double mcol[M];
#pragma unroll
        for (int i = 0; i < M; ++i) {
          s[i] += msk(live, fabs(p - shift[i]));
          s[M + i] += in.J[i] * dy;
          mcol[i] = am * (in.J[i] + lam[i] * eps[i]);
        }
        s[2 * M] += msk(live, p - v) * (p - v);
        s[2 * M + 1] += msk(live, wj) * wj;
        int k = 2 * M + 2;
#pragma unroll
        for (int i = 0; i < M; ++i) {
#pragma unroll
          for (int l = i; l < M; ++l) s[k++] += mcol[i] * mcol[l];
        }
        return (unsigned short)pcode;
      },
      [&](int j, bool live, unsigned short code) {
        if (live) c.pat[j] = code;
      });
  warp_sum_k<NS>(s);
  double Dv = s[2 * M] / 2.0 / d.lr - d.lr / 2.0 * s[2 * M + 1];
#pragma unroll
  for (int i = 0; i < M; ++i) {
    double gi = 0.0;
    if (lsq) gi = P.l1 * s[i];
    else if (L1) gi = P.l1_ratios[i] * s[i];
    const double ci = d.use_c ? d.c[i] : 0.0;
    out.G[i] = gi + s[M + i] + ci;
    Dv += w[i] * (gi + ci);
  }
  out.D = Dv;
  int k = 2 * M + 2;
#pragma unroll
  for (int i = 0; i < M; ++i) {
#pragma unroll
    for (int l = i; l < M; ++l) {
      const double q = d.lr * s[k++];
      out.Q[i][l] = q;
      out.Q[l][i] = q;
    }
  }
}

constexpr __host__ __device__ int zf_popcount(int v) {
  int c = 0;
  while (v) { c += v & 1; v >>= 1; }
  return c;
}
constexpr __host__ __device__ int zf_nth_bit(int mask, int r) {
  int seen = 0;
  for (int b = 0; b < 8; ++b) {
    if (mask & (1 << b)) {
      if (seen == r) return b;
      ++seen;
    }
  }
  return -1;
}

// One face of the simplex (support = MASK): stationary point of the local model
//   G.d - 1/2 d^T Q d,  d = w' - w,
// restricted to the face's affine hull via w'_S = e_0 + Z z, LDL^T on the (K-1)x(K-1)
// reduced PSD matrix.  Everything is expressed in the step d: G is O(1) while Q w can
// reach 1e12 (FDS, n = 100), so G + Q w must never be formed.  A (near-)singular
// reduced matrix means the restricted optimum is not unique or lies on the face's
// boundary, i.e. on a smaller face that is enumerated anyway -- skip.
template <int M, int MASK>
__device__ __forceinline__ void qp_face(const double (&Q)[M][M], const double (&G)[M],
                                        const double (&wc)[M], double pivot_floor,
                                        double (&best_w)[M], double& best_val) {
  constexpr int K = zf_popcount(MASK);
  constexpr int S0 = zf_nth_bit(MASK, 0);
  double w[M];
  bool feas = true;
#pragma unroll
  for (int i = 0; i < M; ++i) w[i] = 0.0;
  if constexpr (K == 1) {
    w[S0] = 1.0;
  } else {
    constexpr int R = K - 1;
    // u = Q (e_S0 - w)
    double u[M];
#pragma unroll
    for (int i = 0; i < M; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int l = 0; l < M; ++l) acc += Q[i][l] * ((l == S0 ? 1.0 : 0.0) - wc[l]);
      u[i] = acc;
    }
    double A[R][R];
    double rhs[R];
#pragma unroll
    for (int a = 0; a < R; ++a) {
      const int ia = zf_nth_bit(MASK, a + 1);
      rhs[a] = (G[ia] - u[ia]) - (G[S0] - u[S0]);
#pragma unroll
      for (int b = 0; b < R; ++b) {
        const int ib = zf_nth_bit(MASK, b + 1);
        A[a][b] = Q[ia][ib] - Q[ia][S0] - Q[S0][ib] + Q[S0][S0];
      }
    }
    // branch free (a rejected face still runs to the end, its result is discarded): the
    // 2^M - 1 faces are independent, so straight-line code lets them overlap
    bool ok = true;
    double inv_piv[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const double piv = A[k][k];
      ok = ok && (piv > pivot_floor);
#if !ZF_QP_BRANCHFREE
      if (!ok) return;
#endif
      const double inv = 1.0 / (ok ? piv : 1.0);
      inv_piv[k] = inv;
#pragma unroll
      for (int r = k + 1; r < R; ++r) {
        const double fct = A[r][k] * inv;
#pragma unroll
        for (int cc = k + 1; cc < R; ++cc) A[r][cc] -= fct * A[k][cc];
        rhs[r] -= fct * rhs[k];
      }
    }
    double z[R];
#pragma unroll
    for (int k = R - 1; k >= 0; --k) {
      double acc = rhs[k];
#pragma unroll
      for (int cc = k + 1; cc < R; ++cc) acc -= A[k][cc] * z[cc];
      z[k] = acc * inv_piv[k];     // the reciprocal of the pivot is already there
    }
    double zs = 0.0;
#pragma unroll
    for (int a = 0; a < R; ++a) {
      zs += z[a];
      feas = feas && (z[a] >= 0.0);
      w[zf_nth_bit(MASK, a + 1)] = z[a];
    }
    const double w0 = 1.0 - zs;
    feas = feas && ok && (w0 >= 0.0);
#if !ZF_QP_BRANCHFREE
    if (!feas) return;
#endif
    w[S0] = w0;
  }
  double dvec[M];
#pragma unroll
  for (int i = 0; i < M; ++i) dvec[i] = w[i] - wc[i];
  double val = 0.0;
#pragma unroll
  for (int i = 0; i < M; ++i) {
    double qd = 0.0;
#pragma unroll
    for (int l = 0; l < M; ++l) qd += Q[i][l] * dvec[l];
    val += dvec[i] * (G[i] - 0.5 * qd);
  }
  if (feas && val > best_val) {
    best_val = val;
#pragma unroll
    for (int i = 0; i < M; ++i) best_w[i] = w[i];
  }
}

template <int M, int MASK>
struct QpFaces {
  __device__ __forceinline__ static void run(const double (&Q)[M][M], const double (&G)[M],
                                             const double (&wc)[M], double pf, double (&bw)[M],
                                             double& bv) {
    qp_face<M, MASK>(Q, G, wc, pf, bw, bv);
    // The first face is the whole simplex: if the maximiser over its affine hull is feasible
    // it is the global maximiser of the concave model and the sub-faces need not be visited.
    if (MASK == (1 << M) - 1 && bv > -CUDART_INF) return;
    QpFaces<M, MASK - 1>::run(Q, G, wc, pf, bw, bv);
  }
};
template <int M>
struct QpFaces<M, 0> {
  __device__ __forceinline__ static void run(const double (&)[M][M], const double (&)[M],
                                             const double (&)[M], double, double (&)[M],
                                             double&) {}
};

// w_out = argmax over the simplex of the local model around wc
template <int M>
__device__ void simplex_qp(const double (&Q)[M][M], const double (&G)[M],
                           const double (&wc)[M], double (&w_out)[M]) {
  double tr = 0.0;
#pragma unroll
  for (int i = 0; i < M; ++i) tr += Q[i][i];
  const double pivot_floor = 1e-13 * tr;
  double best_val = -CUDART_INF;
#pragma unroll
  for (int i = 0; i < M; ++i) w_out[i] = wc[i];
  QpFaces<M, (1 << M) - 1>::run(Q, G, wc, pivot_floor, w_out, best_val);
}

// Returns D(w*); w is in/out (initial guess -> maximiser); *nfev dual evaluations.
// Each step solves the QP of the current quadratic piece exactly.  When the model's
// predicted gain drops below the rounding level of D the step is taken on trust and
// the iteration stops (oracle/dual_model.py:simplex_newton is the CPU statement).
// `probe(wn)` must write x = prox_wsum_g(lr * wn, y - lr * wn @ J) to c.xn and return whether
// every coordinate is on the piece stored in c.pat (primal_probe, or the batched kernel's fused
// sweep that also evaluates F(x) and max|x - y| while it is there).
template <int KIND, int M, int GF, class Probe>
__device__ double dual_newton(const zf_problem& P, const WarpCtx& c, const DualData<M>& d,
                              double (&w)[M], int max_iter, int* nfev, bool* x_ready,
                              Probe&& probe) {
  // One evaluation site and one QP site (code size: see zf_common.cuh): the loop body is
  // "evaluate the point under test, then either accept it and take the next Newton step from
  // it, or halve the step".
  DualPoint<M> cur, pt;
  double wt[M], dir[M];
#pragma unroll
  for (int i = 0; i < M; ++i) { wt[i] = w[i]; dir[i] = 0.0; }
  double step = 1.0;
  int evals = 0, it = 0, bt = 0;
  bool first = true;
  for (;;) {
    dual_full<KIND, M, GF>(P, c, d, wt, pt);
    ++evals;
    if (first || pt.D >= cur.D) {
      first = false;
#pragma unroll
      for (int i = 0; i < M; ++i) w[i] = wt[i];
      cur = pt;
    } else {
      step *= 0.5;
      if (++bt >= 30) break;
#pragma unroll
      for (int i = 0; i < M; ++i) wt[i] = w[i] + step * dir[i];
      continue;
    }
    if (++it > max_iter) break;
    double gabs = 0.0, wn[M];
#pragma unroll
    for (int i = 0; i < M; ++i) gabs = fmax(gabs, fabs(cur.G[i]));
    simplex_qp<M>(cur.Q, cur.G, w, wn);
    double dmax = 0.0;
#pragma unroll
    for (int i = 0; i < M; ++i) {
      dir[i] = wn[i] - w[i];
      dmax = fmax(dmax, fabs(dir[i]));
    }
    if (dmax == 0.0) break;
    double pred = 0.0;
#pragma unroll
    for (int i = 0; i < M; ++i) {
      double qd = 0.0;
#pragma unroll
      for (int l = 0; l < M; ++l) qd += cur.Q[i][l] * dir[l];
      pred += dir[i] * (cur.G[i] - 0.5 * qd);
    }
    if (pred <= 1e-15 * (fabs(cur.D) + gabs)) {
      // below the rounding level of D: take the Newton step on trust and stop
#pragma unroll
      for (int i = 0; i < M; ++i) w[i] = wn[i];
      cur.D += pred;
      break;
    }
    // The dual is one quadratic as long as no coordinate changes piece.  If the primal point
    // of the Newton candidate lies on the same pieces as the point just evaluated, the model
    // is exact there: the candidate is the maximiser, D(candidate) = D + pred, and evaluating
    // it (plus the QP that would return it unchanged) is only a confirmation -- skip both.
    // The probe is the primal recovery the caller needs anyway, so it costs nothing extra.
    if (probe(wn)) {
#pragma unroll
      for (int i = 0; i < M; ++i) w[i] = wn[i];
      cur.D += pred;
      *x_ready = true;
      break;
    }
    step = 1.0;
    bt = 0;
#pragma unroll
    for (int i = 0; i < M; ++i) wt[i] = wn[i];
  }
  *nfev = evals;
  return cur.D;
}

}  // namespace zf
