// Device functors for the problem classes of zfista/problems.py.
//
// Each Fn<KIND, M> gives warp-collective evaluations with the semantics of the
// reference methods:
//     f(x)            -> Problem.f          (returns the M objective values)
//     f_jac(y, J)     -> Problem.f + Problem.jac_f  (J rows written to shared memory)
// g() and the prox chain below are the shared Problem.g / Problem.prox_wsum_g
// (problems.py:101-138).  All 32 lanes return identical values.
#pragma once
#include "zf_common.cuh"

namespace zf {

template <int KIND, int M>
struct Fn;

// ---------------------------------------------------------------- JOS1 (problems.py:153-205)
template <>
struct Fn<ZF_JOS1, 2> {
  __device__ static void f(const zf_problem& P, const WarpCtx& c, const double* x,
                           double (&out)[2]) {
    double s[2] = {0.0, 0.0};
#pragma unroll 1
    for (int j = c.lane; j < c.n; j += 32) {
      const double xj = x[j];
      s[0] += xj * xj;
      s[1] += (xj - 2.0) * (xj - 2.0);
    }
    warp_sum_k<2>(s);
    out[0] = norm_sq_like_numpy(s[0]) / (double)c.n;
    out[1] = norm_sq_like_numpy(s[1]) / (double)c.n;
  }
  __device__ static void f_jac(const zf_problem& P, const WarpCtx& c, const double* y,
                               double* J, double (&fy)[2]) {
    f(P, c, y, fy);
    const double dn = (double)c.n;
#pragma unroll 1
    for (int j = c.lane; j < c.n; j += 32) {
      const double yj = y[j];
      J[j] = 2.0 * yj / dn;
      J[c.n + j] = 2.0 * (yj - 2.0) / dn;
    }
  }
};

// ---------------------------------------------------------------- SD (problems.py:208-264)
template <>
struct Fn<ZF_SD, 2> {
  __device__ static void f(const zf_problem& P, const WarpCtx& c, const double* x,
                           double (&out)[2]) {
    const double r2 = sqrt(2.0);
    const double x0 = x[0], x1 = x[1], x2 = x[2], x3 = x[3];
    out[0] = 2.0 * x0 + r2 * x1 + r2 * x2 + x3;
    out[1] = 2.0 / x0 + 2.0 * r2 / x1 + 2.0 * r2 / x2 + 2.0 / x3;
  }
  __device__ static void f_jac(const zf_problem& P, const WarpCtx& c, const double* y,
                               double* J, double (&fy)[2]) {
    f(P, c, y, fy);
    const double r2 = sqrt(2.0);
    if (c.lane < 4) {
      const int j = c.lane;
      const double yj = y[j];
      const double a = (j == 0) ? 2.0 : (j == 3 ? 1.0 : r2);
      const double bnum = (j == 0 || j == 3) ? -2.0 : -2.0 * r2;
      J[j] = a;
      J[4 + j] = bnum / (yj * yj);
    }
  }
};

// ---------------------------------------------------------------- FDS (problems.py:267-328)
template <>
struct Fn<ZF_FDS, 3> {
  __device__ static void moments(const WarpCtx& c, const double* x, double (&s)[4]) {
    s[0] = s[1] = s[2] = s[3] = 0.0;
#pragma unroll 1
    for (int j = c.lane; j < c.n; j += 32) {
      const double xj = x[j];
      const double idx = (double)(j + 1);
      const double d = xj - idx;
      const double d2 = d * d;
      s[0] += idx * (d2 * d2);
      s[1] += xj;
      s[2] += xj * xj;
      s[3] += (idx * (double)(c.n - j)) * exp(-xj);
    }
    warp_sum_k<4>(s);
  }
  __device__ static void finish(const WarpCtx& c, const double (&s)[4], double (&out)[3],
                                double& e_mean) {
    const double dn = (double)c.n;
    e_mean = exp(s[1] / dn);
    out[0] = s[0] / (dn * dn);
    out[1] = e_mean + norm_sq_like_numpy(s[2]);
    out[2] = s[3] / (dn * (dn + 1.0));
  }
  __device__ static void f(const zf_problem& P, const WarpCtx& c, const double* x,
                           double (&out)[3]) {
    double s[4], e;
    moments(c, x, s);
    finish(c, s, out, e);
  }
  __device__ static void f_jac(const zf_problem& P, const WarpCtx& c, const double* y,
                               double* J, double (&fy)[3]) {
    // one pass: the moments of f and the Jacobian rows 1 and 3 share exp(-y_j); row 2 needs
    // exp(mean(y)), known only after the reduction, and is filled by a second (cheap) pass
    double s[4], e;
    s[0] = s[1] = s[2] = s[3] = 0.0;
    const double dn = (double)c.n;
    const double c1 = 4.0 / (dn * dn);
    const double c3 = dn * (dn + 1.0);
#pragma unroll 1
    for (int j = c.lane; j < c.n; j += 32) {
      const double yj = y[j];
      const double idx = (double)(j + 1);
      const double d = yj - idx;
      const double d2 = d * d;
      const double conv = idx * (double)(c.n - j);
      const double ex = exp(-yj);
      s[0] += idx * (d2 * d2);
      s[1] += yj;
      s[2] += yj * yj;
      s[3] += conv * ex;
      J[j] = c1 * idx * (d * d * d);
      J[2 * c.n + j] = -conv * ex / c3;
    }
    warp_sum_k<4>(s);
    finish(c, s, fy, e);
    const double e_over_n = e / dn;
#pragma unroll 1
    for (int j = c.lane; j < c.n; j += 32) J[c.n + j] = e_over_n + 2.0 * y[j];
  }
};

// ---------------------------------------------------------------- ZDT1 (problems.py:331-386)
template <>
struct Fn<ZF_ZDT1, 2> {
  __device__ static double h_of(const WarpCtx& c, const double* x) {
    double s = 0.0;
#pragma unroll 1
    for (int j = c.lane; j < c.n; j += 32) s += (j >= 1) ? x[j] : 0.0;
    s = warp_sum(s);
    return 1.0 + 9.0 / (double)(c.n - 1) * s;
  }
  __device__ static void f(const zf_problem& P, const WarpCtx& c, const double* x,
                           double (&out)[2]) {
    const double h = h_of(c, x);
    const double x0 = x[0];
    out[0] = x0;
    out[1] = h * (1.0 - sqrt(x0 / h));
  }
  __device__ static void f_jac(const zf_problem& P, const WarpCtx& c, const double* y,
                               double* J, double (&fy)[2]) {
    const double h = h_of(c, y);
    const double y0 = y[0];
    fy[0] = y0;
    fy[1] = h * (1.0 - sqrt(y0 / h));
    const double rest = 9.0 * (2.0 - sqrt(y0 / h)) / 2.0 / (double)(c.n - 1);
    const double first = -sqrt(h / y0) / 2.0;
#pragma unroll 1
    for (int j = c.lane; j < c.n; j += 32) {
      J[j] = (j == 0) ? 1.0 : 0.0;
      J[c.n + j] = (j == 0) ? first : rest;
    }
  }
};

// ---------------------------------------------------------------- TOI4 (problems.py:389-448)
template <>
struct Fn<ZF_TOI4, 2> {
  __device__ static void f(const zf_problem& P, const WarpCtx& c, const double* x,
                           double (&out)[2]) {
    const double x0 = x[0], x1 = x[1], x2 = x[2], x3 = x[3];
    out[0] = x0 * x0 + x1 * x1 + 1.0;
    out[1] = 0.5 * ((x0 - x1) * (x0 - x1) + (x2 - x3) * (x2 - x3)) + 1.0;
  }
  __device__ static void f_jac(const zf_problem& P, const WarpCtx& c, const double* y,
                               double* J, double (&fy)[2]) {
    f(P, c, y, fy);
    if (c.lane == 0) {
      const double y0 = y[0], y1 = y[1], y2 = y[2], y3 = y[3];
      J[0] = 2.0 * y0;
      J[1] = 2.0 * y1;
      J[2] = 0.0;
      J[3] = 0.0;
      const double a = y0 - y1, b = y2 - y3;
      J[4] = a;
      J[5] = -a;
      J[6] = b;
      J[7] = -b;
    }
  }
};

// ---------------------------------------------------------------- TRIDIA (problems.py:451-514)
template <>
struct Fn<ZF_TRIDIA, 3> {
  __device__ static void f(const zf_problem& P, const WarpCtx& c, const double* x,
                           double (&out)[3]) {
    const double x0 = x[0], x1 = x[1], x2 = x[2];
    out[0] = sq(2.0 * x0 - 1.0);
    out[1] = 2.0 * sq(2.0 * x0 - x1);
    out[2] = 3.0 * sq(2.0 * x1 - x2);
  }
  __device__ static void f_jac(const zf_problem& P, const WarpCtx& c, const double* y,
                               double* J, double (&fy)[3]) {
    f(P, c, y, fy);
    if (c.lane == 0) {
      const double y0 = y[0], y1 = y[1], y2 = y[2];
      J[0] = 8.0 * y0 - 4.0;
      J[1] = 0.0;
      J[2] = 0.0;
      J[3] = 16.0 * y0 - 8.0 * y1;
      J[4] = 4.0 * y1 - 8.0 * y0;
      J[5] = 0.0;
      J[6] = 0.0;
      J[7] = 24.0 * y1 - 12.0 * y2;
      J[8] = 6.0 * y2 - 12.0 * y1;
    }
  }
};

// ------------------------------------------------- LinearFunctionRank1 (problems.py:517-578)
template <int M>
struct Fn<ZF_LFR1, M> {
  __device__ static double weighted_sum(const WarpCtx& c, const double* x) {
    double s = 0.0;
#pragma unroll 1
    for (int j = c.lane; j < c.n; j += 32) s += (double)(j + 1) * x[j];
    return warp_sum(s);
  }
  __device__ static void f(const zf_problem& P, const WarpCtx& c, const double* x,
                           double (&out)[M]) {
    const double s = weighted_sum(c, x);
#pragma unroll
    for (int i = 0; i < M; ++i) out[i] = sq((double)(i + 1) * s - 1.0);
  }
  __device__ static void f_jac(const zf_problem& P, const WarpCtx& c, const double* y,
                               double* J, double (&fy)[M]) {
    const double s = weighted_sum(c, y);
#pragma unroll
    for (int i = 0; i < M; ++i) fy[i] = sq((double)(i + 1) * s - 1.0);
#pragma unroll 1
    for (int j = c.lane; j < c.n; j += 32) {
#pragma unroll
      for (int i = 0; i < M; ++i) {
        const double di = (double)(i + 1);
        // 2 * i * k * (i * s - 1), evaluated left to right like numpy broadcasting
        J[i * c.n + j] = 2.0 * di * (double)(j + 1) * (di * s - 1.0);
      }
    }
  }
};

// --------------------------- scale*||Ax-b||^2 replicated M times (tests/test_proximal_gradient.py)
template <int M>
struct Fn<ZF_LSQ_L1, M> {
  // r = A x - b into scratch; returns scale * ||r||^2
  __device__ static double residual(const zf_problem& P, const WarpCtx& c, const double* x) {
    double ss = 0.0;
    for (int r = 0; r < P.n_rows; ++r) {
      const double* Ar = P.A + (size_t)r * c.n;
      double d = 0.0;
#pragma unroll 1
      for (int j = c.lane; j < c.n; j += 32) d += Ar[j] * x[j];
      d = warp_sum(d) - P.b[r];
      if (c.lane == 0) c.scratch[r] = d;
      ss += d * d;
    }
    __syncwarp();
    return norm_sq_like_numpy(ss) * P.scale;
  }
  __device__ static void f(const zf_problem& P, const WarpCtx& c, const double* x,
                           double (&out)[M]) {
    const double v = residual(P, c, x);
#pragma unroll
    for (int i = 0; i < M; ++i) out[i] = v;
  }
  __device__ static void f_jac(const zf_problem& P, const WarpCtx& c, const double* y,
                               double* J, double (&fy)[M]) {
    const double v = residual(P, c, y);
#pragma unroll
    for (int i = 0; i < M; ++i) fy[i] = v;
    const double two_scale = 2.0 * P.scale;
#pragma unroll 1
    for (int j = c.lane; j < c.n; j += 32) {
      double gsum = 0.0;
      for (int r = 0; r < P.n_rows; ++r) gsum += P.A[(size_t)r * c.n + j] * c.scratch[r];
      gsum *= two_scale;
#pragma unroll
      for (int i = 0; i < M; ++i) J[i * c.n + j] = gsum;
    }
    __syncwarp();
  }
};

// ======================================================================================
// g and prox_wsum_g  (problems.py:101-138)
// ======================================================================================
__device__ __forceinline__ double lower_of(const zf_problem& P, int j) {
  return P.bounds_are_arrays ? P.lower_v[j] : P.lower;
}
__device__ __forceinline__ double upper_of(const zf_problem& P, int j) {
  return P.bounds_are_arrays ? P.upper_v[j] : P.upper;
}

// Problem.g: +inf for every objective outside the box, else l1_ratios_i*||x - shift_i||_1
template <int M>
__device__ void g_eval(const zf_problem& P, const WarpCtx& c, const double* x,
                       double (&out)[M]) {
  if (P.kind == ZF_LSQ_L1) {
    double s = 0.0;
#pragma unroll 1
    for (int j = c.lane; j < c.n; j += 32) s += fabs(x[j]);
    s = warp_sum(s) * P.l1;
#pragma unroll
    for (int i = 0; i < M; ++i) out[i] = s;
    return;
  }
  if (P.has_bounds) {
    int bad = 0;
#pragma unroll 1
    for (int j = c.lane; j < c.n; j += 32) {
      const double xj = x[j];
      bad |= (xj < lower_of(P, j)) || (xj > upper_of(P, j));
    }
    if (__any_sync(ZF_FULL_MASK, bad)) {
#pragma unroll
      for (int i = 0; i < M; ++i) out[i] = CUDART_INF;
      return;
    }
  }
  if (P.has_l1) {
    double s[M];
#pragma unroll
    for (int i = 0; i < M; ++i) s[i] = 0.0;
#pragma unroll 1
    for (int j = c.lane; j < c.n; j += 32) {
      const double xj = x[j];
#pragma unroll
      for (int i = 0; i < M; ++i) s[i] += fabs(xj - P.l1_shifts[i]);
    }
    warp_sum_k<M>(s);
#pragma unroll
    for (int i = 0; i < M; ++i) out[i] = P.l1_ratios[i] * s[i];
    return;
  }
#pragma unroll
  for (int i = 0; i < M; ++i) out[i] = 0.0;
}

// Per-coordinate prox with the reference's stage order.  `wt[i]` is the weight
// argument of prox_wsum_g (the caller passes lr * w_i).  With TRACK the function
// also reports whether the coordinate is free (alpha = 1) or pinned at a kink /
// bound (alpha = 0) and on which side of shift i it sits (eps[i] = +-1); the
// simplex-Newton dual solver builds its generalised Hessian from these.
template <int M, bool TRACK>
__device__ __forceinline__ double prox_elem(const zf_problem& P, int j, double v,
                                            const double (&wt)[M], double& alpha,
                                            double (&eps)[M]) {
  double p = v;
  if (TRACK) {
    alpha = 1.0;
#pragma unroll
    for (int i = 0; i < M; ++i) eps[i] = 0.0;
  }
  if (P.kind == ZF_LSQ_L1) {
    double wsum = 0.0;
#pragma unroll
    for (int i = 0; i < M; ++i) wsum += wt[i];
    const double t = P.l1 * wsum;
    p = soft_threshold(v, t);
    if (TRACK) {
      if (fabs(v) <= t) alpha = 0.0;
#pragma unroll
      for (int i = 0; i < M; ++i) eps[i] = v > t ? 1.0 : -1.0;
    }
    return p;
  }
  if (P.has_l1) {
    double coef[M];
#pragma unroll
    for (int i = 0; i < M; ++i) coef[i] = wt[i] * P.l1_ratios[i];
    double tail = 0.0;
#pragma unroll
    for (int i = 1; i < M; ++i) tail += coef[i];
    const double a0 = p + tail - P.l1_shifts[0] + P.l1_shifts[0];
    p = soft_threshold(a0, coef[0]);
    if (TRACK) {
      if (fabs(a0) <= coef[0]) alpha = 0.0;
      eps[0] = a0 > coef[0] ? 1.0 : -1.0;
    }
#pragma unroll
    for (int i = 1; i < M; ++i) {
      const double ai = p - coef[i] - P.l1_shifts[i];
      p = soft_threshold(ai, coef[i]) + P.l1_shifts[i];
      if (TRACK) {
        if (fabs(ai) <= coef[i]) alpha = 0.0;
        eps[i] = ai > coef[i] ? 1.0 : -1.0;
      }
    }
  }
  if (P.has_bounds) {
    const double lo = lower_of(P, j), hi = upper_of(P, j);
    const double q = fmin(fmax(p, lo), hi);
    if (TRACK) {
      if (q != p) alpha = 0.0;
    }
    p = q;
  }
  return p;
}

}  // namespace zf
