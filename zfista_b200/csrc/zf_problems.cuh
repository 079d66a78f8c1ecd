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

// parts of g a kernel instantiation handles (template flag GF)
constexpr int ZF_G_L1 = 1, ZF_G_BOX = 2;

template <int KIND, int M>
struct Fn;

// Where f_jac reads y^k from.  YPlain: the vector is in shared memory.  YFold: the momentum
// extrapolation of the previous iteration (proximal_gradient.py:530-538) is folded into this
// read -- y_j = x_j + mom (x_j - xold_j) (FISTA) or y_j = x_j (ISTA) is formed on the fly and
// written to y, which saves the separate extrapolation sweep over the coordinates.
struct YPlain {
  const double* y;
  __device__ __forceinline__ double load(int jc) const { return y[jc]; }
  __device__ __forceinline__ void store(int, bool, double) const {}
};
struct YFold {
  double* y;
  const double* xnew;
  const double* xold;
  double mom;
  bool nesterov;
  __device__ __forceinline__ double load(int jc) const {
    const double xj = xnew[jc];
    const double ext = xj + mom * (xj - xold[jc]);     // both operands always readable: no branch
    return nesterov ? ext : xj;
  }
  __device__ __forceinline__ void store(int j, bool live, double v) const {
    if (live) y[j] = v;
  }
};

// Every functor below states f as  finish(sum_j terms_j):  f_pre computes a coordinate's terms,
// f_acc adds them (masked for dead sweep slots) to the lane's partial sums, f_finish turns the
// warp totals into the M objective values.  That split is what lets the batched kernel evaluate
// F(x) inside the sweep that forms x (fused_primal_eval in zf_batched_kernels.cuh).
// kPointwise = false: f is not such a sum (dense A).  kFoldable: f_jac takes its input through a
// YFold.  FAST = the branch-free forms of exp / division (zf_common.cuh); a sweep that raised
// `rare` is re-run with FAST = false.

// ---------------------------------------------------------------- JOS1 (problems.py:153-205)
template <>
struct Fn<ZF_JOS1, 2> {
  static constexpr bool kPointwise = true, kFoldable = true;
  static constexpr int NF = 2, NT = 2;
  // divisor shared by every coordinate and iteration: its reciprocal is refined once per start
  struct Consts { Recip dn; };
  __device__ __forceinline__ static Consts make_consts(const WarpCtx& c) {
    return Consts{make_recip((double)c.n)};
  }
  template <bool FAST>
  __device__ __forceinline__ static void f_pre(const WarpCtx& c, int j, double xj,
                                               double (&t)[NT], int& rare) {
    t[0] = xj;
    t[1] = xj - 2.0;
  }
  __device__ __forceinline__ static void f_acc(bool live, const double (&t)[NT],
                                               double (&s)[NF]) {
    s[0] += msk(live, t[0]) * t[0];
    s[1] += msk(live, t[1]) * t[1];
  }
  __device__ __forceinline__ static void f_finish(const zf_problem& P, const WarpCtx& c,
                                                  const Consts& K, const double* x, const double (&s)[NF],
                                                  double (&out)[2]) {
    out[0] = div_exact(norm_sq_like_numpy(s[0]), K.dn);
    out[1] = div_exact(norm_sq_like_numpy(s[1]), K.dn);
  }
  struct JacOut { double y, j0, j1; };
  template <bool FAST, class YL>
  __device__ __forceinline__ static int jac_sweep(const WarpCtx& c, const YL& Y, double* J,
                                                  double (&s)[NF], const Recip& dnr) {
    int rare = 0;
    sweep3<double, JacOut>(
        c.n, c.lane,
        [&](int j, bool live) { return Y.load(live ? j : 0); },
        [&](double yj, int j, bool live) {
          JacOut o;
          o.y = yj;
          o.j0 = FAST ? div_regular(2.0 * yj, dnr, rare) : 2.0 * yj / dnr.c;
          o.j1 = FAST ? div_regular(2.0 * (yj - 2.0), dnr, rare) : 2.0 * (yj - 2.0) / dnr.c;
          s[0] += msk(live, yj) * yj;
          s[1] += msk(live, yj - 2.0) * (yj - 2.0);
          return o;
        },
        [&](int j, bool live, const JacOut& o) {
          Y.store(j, live, o.y);
          if (live) {
            J[j] = o.j0;
            J[c.n + j] = o.j1;
          }
        });
    return rare;
  }
  template <class YL>
  __device__ static void f_jac(const zf_problem& P, const WarpCtx& c, const Consts& K, const YL& Y,
                               double* J,
                               double (&fy)[2]) {
    double s[NF] = {0.0, 0.0};
    if (__any_sync(ZF_FULL_MASK, jac_sweep<true>(c, Y, J, s, K.dn))) {
      s[0] = s[1] = 0.0;
      jac_sweep<false>(c, Y, J, s, K.dn);
    }
    warp_sum_k<NF>(s);
    f_finish(P, c, K, nullptr, s, fy);
  }
};

// ---------------------------------------------------------------- SD (problems.py:208-264)
template <>
struct Fn<ZF_SD, 2> {
  static constexpr bool kPointwise = true, kFoldable = false;
  static constexpr int NF = 1, NT = 1;
  struct Consts {};
  __device__ __forceinline__ static Consts make_consts(const WarpCtx&) { return Consts{}; }
  template <bool FAST>
  __device__ __forceinline__ static void f_pre(const WarpCtx&, int, double, double (&)[NT], int&) {}
  __device__ __forceinline__ static void f_acc(bool, const double (&)[NT], double (&)[NF]) {}
  __device__ __forceinline__ static void f_finish(const zf_problem& P, const WarpCtx& c,
                                                  const Consts& K, const double* x, const double (&)[NF],
                                                  double (&out)[2]) {
    const double r2 = sqrt(2.0);
    const double x0 = x[0], x1 = x[1], x2 = x[2], x3 = x[3];
    out[0] = 2.0 * x0 + r2 * x1 + r2 * x2 + x3;
    out[1] = 2.0 / x0 + 2.0 * r2 / x1 + 2.0 * r2 / x2 + 2.0 / x3;
  }
  __device__ static void f_jac(const zf_problem& P, const WarpCtx& c, const Consts& K, const YPlain& Y,
                               double* J,
                               double (&fy)[2]) {
    const double* y = Y.y;
    double none[NF] = {0.0};
    f_finish(P, c, K, y, none, fy);
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
  static constexpr bool kPointwise = true, kFoldable = true;
  static constexpr int NF = 4, NT = 5;
  // divisors shared by every coordinate and iteration: reciprocals refined once per start
  struct Consts { Recip dn, dn2, c3; double c1; };
  __device__ __forceinline__ static Consts make_consts(const WarpCtx& c) {
    const double dn = (double)c.n;
    return Consts{make_recip(dn), make_recip(dn * dn), make_recip(dn * (dn + 1.0)),
                  4.0 / (dn * dn)};
  }
  template <bool FAST>
  __device__ __forceinline__ static void f_pre(const WarpCtx& c, int j, double xj,
                                               double (&t)[NT], int& rare) {
    const double idx = (double)(j + 1);
    const double d = xj - idx;
    const double d2 = d * d;
    t[0] = idx;
    t[1] = d2 * d2;
    t[2] = xj;
    t[3] = idx * (double)(c.n - j);
    t[4] = FAST ? exp_regular(-xj, rare) : exp(-xj);
  }
  __device__ __forceinline__ static void f_acc(bool live, const double (&t)[NT],
                                               double (&s)[NF]) {
    s[0] += msk(live, t[0]) * t[1];
    s[1] += msk(live, t[2]);
    s[2] += msk(live, t[2]) * t[2];
    s[3] += msk(live, t[3]) * t[4];
  }
  __device__ __forceinline__ static void finish(const WarpCtx& c, const Consts& K,
                                                const double (&s)[4], double (&out)[3],
                                                double& e_mean) {
    e_mean = exp_exact(div_exact(s[1], K.dn));
    out[0] = div_exact(s[0], K.dn2);
    out[1] = e_mean + norm_sq_like_numpy(s[2]);
    out[2] = div_exact(s[3], K.c3);
  }
  __device__ __forceinline__ static void f_finish(const zf_problem& P, const WarpCtx& c,
                                                  const Consts& K, const double* x,
                                                  const double (&s)[NF], double (&out)[3]) {
    double e;
    finish(c, K, s, out, e);
  }
  struct JacOut { double y, j0, j2; };
  template <bool FAST, class YL>
  __device__ __forceinline__ static int jac_sweep(const WarpCtx& c, const YL& Y, double* J,
                                                  double (&s)[NF], double c1, const Recip& c3r) {
    int rare = 0;
    sweep3<double, JacOut>(
        c.n, c.lane,
        [&](int j, bool live) { return Y.load(live ? j : 0); },
        [&](double yj, int j, bool live) {
          const int jc = live ? j : 0;
          const double idx = (double)(jc + 1);
          const double d = yj - idx;
          const double d2 = d * d;
          const double conv = idx * (double)(c.n - jc);
          const double ex = FAST ? exp_regular(-yj, rare) : exp(-yj);
          JacOut o;
          o.y = yj;
          o.j0 = c1 * idx * (d * d * d);
          o.j2 = FAST ? div_regular(-conv * ex, c3r, rare) : -conv * ex / c3r.c;
          s[0] += msk(live, idx) * (d2 * d2);
          s[1] += msk(live, yj);
          s[2] += msk(live, yj) * yj;
          s[3] += msk(live, conv) * ex;
          return o;
        },
        [&](int j, bool live, const JacOut& o) {
          Y.store(j, live, o.y);
          if (live) {
            J[j] = o.j0;
            J[2 * c.n + j] = o.j2;
          }
        });
    return rare;
  }
  template <class YL>
  __device__ static void f_jac(const zf_problem& P, const WarpCtx& c, const Consts& K, const YL& Y,
                               double* J,
                               double (&fy)[3]) {
    // one sweep: the moments of f and the Jacobian rows 1 and 3 share exp(-y_j); row 2 needs
    // exp(mean(y)), known only after the reduction, and is filled by a second (cheap) sweep
    double s[4] = {0.0, 0.0, 0.0, 0.0}, e;
    if (__any_sync(ZF_FULL_MASK, jac_sweep<true>(c, Y, J, s, K.c1, K.c3))) {
      s[0] = s[1] = s[2] = s[3] = 0.0;
      jac_sweep<false>(c, Y, J, s, K.c1, K.c3);
    }
    warp_sum_k<4>(s);
    finish(c, K, s, fy, e);
    const double e_over_n = div_exact(e, K.dn);
    const double* y = Y.y;        // a lane re-reads only coordinates it wrote itself
    sweep3<double, double>(
        c.n, c.lane, [&](int j, bool live) { return y[live ? j : 0]; },
        [&](double yj, int, bool) { return e_over_n + 2.0 * yj; },
        [&](int j, bool live, double v) {
          if (live) J[c.n + j] = v;
        });
  }
};

// ---------------------------------------------------------------- ZDT1 (problems.py:331-386)
template <>
struct Fn<ZF_ZDT1, 2> {
  static constexpr bool kPointwise = true, kFoldable = true;
  static constexpr int NF = 1, NT = 1;
  struct Consts {};
  __device__ __forceinline__ static Consts make_consts(const WarpCtx&) { return Consts{}; }
  template <bool FAST>
  __device__ __forceinline__ static void f_pre(const WarpCtx& c, int j, double xj,
                                               double (&t)[NT], int& rare) {
    t[0] = (j >= 1) ? xj : 0.0;
  }
  __device__ __forceinline__ static void f_acc(bool live, const double (&t)[NT],
                                               double (&s)[NF]) {
    s[0] += msk(live, t[0]);
  }
  __device__ __forceinline__ static void f_finish(const zf_problem& P, const WarpCtx& c,
                                                  const Consts& K, const double* x, const double (&s)[NF],
                                                  double (&out)[2]) {
    const double h = 1.0 + 9.0 / (double)(c.n - 1) * s[0];
    const double x0 = x[0];
    out[0] = x0;
    out[1] = h * (1.0 - sqrt(x0 / h));
  }
  template <class YL>
  __device__ static void f_jac(const zf_problem& P, const WarpCtx& c, const Consts& K, const YL& Y,
                               double* J,
                               double (&fy)[2]) {
    double s[NF] = {0.0};
    sweep3<double, double>(
        c.n, c.lane, [&](int j, bool live) { return Y.load(live ? j : 0); },
        [&](double yj, int j, bool live) {
          s[0] += msk(live, (j >= 1) ? yj : 0.0);
          return yj;
        },
        [&](int j, bool live, double yj) { Y.store(j, live, yj); });
    warp_sum_k<NF>(s);
    __syncwarp();
    const double h = 1.0 + 9.0 / (double)(c.n - 1) * s[0];
    const double y0 = Y.y[0];
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
  static constexpr bool kPointwise = true, kFoldable = false;
  static constexpr int NF = 1, NT = 1;
  struct Consts {};
  __device__ __forceinline__ static Consts make_consts(const WarpCtx&) { return Consts{}; }
  template <bool FAST>
  __device__ __forceinline__ static void f_pre(const WarpCtx&, int, double, double (&)[NT], int&) {}
  __device__ __forceinline__ static void f_acc(bool, const double (&)[NT], double (&)[NF]) {}
  __device__ __forceinline__ static void f_finish(const zf_problem& P, const WarpCtx& c,
                                                  const Consts& K, const double* x, const double (&)[NF],
                                                  double (&out)[2]) {
    const double x0 = x[0], x1 = x[1], x2 = x[2], x3 = x[3];
    out[0] = x0 * x0 + x1 * x1 + 1.0;
    out[1] = 0.5 * ((x0 - x1) * (x0 - x1) + (x2 - x3) * (x2 - x3)) + 1.0;
  }
  __device__ static void f_jac(const zf_problem& P, const WarpCtx& c, const Consts& K, const YPlain& Y,
                               double* J,
                               double (&fy)[2]) {
    const double* y = Y.y;
    double none[NF] = {0.0};
    f_finish(P, c, K, y, none, fy);
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
  static constexpr bool kPointwise = true, kFoldable = false;
  static constexpr int NF = 1, NT = 1;
  struct Consts {};
  __device__ __forceinline__ static Consts make_consts(const WarpCtx&) { return Consts{}; }
  template <bool FAST>
  __device__ __forceinline__ static void f_pre(const WarpCtx&, int, double, double (&)[NT], int&) {}
  __device__ __forceinline__ static void f_acc(bool, const double (&)[NT], double (&)[NF]) {}
  __device__ __forceinline__ static void f_finish(const zf_problem& P, const WarpCtx& c,
                                                  const Consts& K, const double* x, const double (&)[NF],
                                                  double (&out)[3]) {
    const double x0 = x[0], x1 = x[1], x2 = x[2];
    out[0] = sq(2.0 * x0 - 1.0);
    out[1] = 2.0 * sq(2.0 * x0 - x1);
    out[2] = 3.0 * sq(2.0 * x1 - x2);
  }
  __device__ static void f_jac(const zf_problem& P, const WarpCtx& c, const Consts& K, const YPlain& Y,
                               double* J,
                               double (&fy)[3]) {
    const double* y = Y.y;
    double none[NF] = {0.0};
    f_finish(P, c, K, y, none, fy);
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
  static constexpr bool kPointwise = true, kFoldable = true;
  static constexpr int NF = 1, NT = 2;
  struct Consts {};
  __device__ __forceinline__ static Consts make_consts(const WarpCtx&) { return Consts{}; }
  template <bool FAST>
  __device__ __forceinline__ static void f_pre(const WarpCtx& c, int j, double xj,
                                               double (&t)[NT], int& rare) {
    t[0] = (double)(j + 1);
    t[1] = xj;
  }
  __device__ __forceinline__ static void f_acc(bool live, const double (&t)[NT],
                                               double (&s)[NF]) {
    s[0] += msk(live, t[0]) * t[1];
  }
  __device__ __forceinline__ static void f_finish(const zf_problem& P, const WarpCtx& c,
                                                  const Consts& K, const double* x, const double (&s)[NF],
                                                  double (&out)[M]) {
#pragma unroll
    for (int i = 0; i < M; ++i) out[i] = sq((double)(i + 1) * s[0] - 1.0);
  }
  template <class YL>
  __device__ static void f_jac(const zf_problem& P, const WarpCtx& c, const Consts& K, const YL& Y,
                               double* J,
                               double (&fy)[M]) {
    double s[NF] = {0.0};
    sweep3<double, double>(
        c.n, c.lane, [&](int j, bool live) { return Y.load(live ? j : 0); },
        [&](double yj, int j, bool live) {
          s[0] += msk(live, (double)(j + 1)) * yj;
          return yj;
        },
        [&](int j, bool live, double yj) { Y.store(j, live, yj); });
    warp_sum_k<NF>(s);
    f_finish(P, c, K, nullptr, s, fy);
    const double sum = s[0];
#pragma unroll 1
    for (int j = c.lane; j < c.n; j += 32) {
#pragma unroll
      for (int i = 0; i < M; ++i) {
        const double di = (double)(i + 1);
        // 2 * i * k * (i * s - 1), evaluated left to right like numpy broadcasting
        J[i * c.n + j] = 2.0 * di * (double)(j + 1) * (di * sum - 1.0);
      }
    }
  }
};

// --------------------------- scale*||Ax-b||^2 replicated M times (tests/test_proximal_gradient.py)
template <int M>
struct Fn<ZF_LSQ_L1, M> {
  static constexpr bool kPointwise = false, kFoldable = false;
  static constexpr int NF = 1, NT = 1;
  struct Consts {};
  __device__ __forceinline__ static Consts make_consts(const WarpCtx&) { return Consts{}; }
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
  __device__ static void f_jac(const zf_problem& P, const WarpCtx& c, const Consts& K, const YPlain& Y,
                               double* J,
                               double (&fy)[M]) {
    const double v = residual(P, c, Y.y);
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

// Problem.f at a vector in shared memory (all lanes return the same values)
template <int KIND, int M>
__device__ void f_eval(const zf_problem& P, const WarpCtx& c,
                       const typename Fn<KIND, M>::Consts& K, const double* x, double (&out)[M]) {
  using F = Fn<KIND, M>;
  if constexpr (!F::kPointwise) {
    F::f(P, c, x, out);
  } else {
    double s[F::NF];
#pragma unroll
    for (int k = 0; k < F::NF; ++k) s[k] = 0.0;
    if constexpr (F::kFoldable) {       // (the fixed-size problems read x[0..3] in f_finish)
      int rare = 0;
      sweep(c.n, c.lane, [&](int j, bool live) {
        const int jc = live ? j : 0;
        double t[F::NT];
        F::template f_pre<false>(c, jc, x[jc], t, rare);
        F::f_acc(live, t, s);
      });
      warp_sum_k<F::NF>(s);
    }
    F::f_finish(P, c, K, x, s, out);
  }
}

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
template <int KIND, int M>
__device__ void g_eval(const zf_problem& P, const WarpCtx& c, const double* x,
                       double (&out)[M]) {
  if constexpr (KIND == ZF_LSQ_L1) {
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

// Per-coordinate prox with the reference's stage order (problems.py:126-137).  `wt[i]` is the
// weight argument of prox_wsum_g (the caller passes lr * w_i).  GF = which parts g has: ZF_G_L1
// (l1_ratios given) | ZF_G_BOX (bounds given) -- template flags: the kernels are instantiated per
// combination, so that the hot sweeps carry neither a branch nor dead work for an absent part.  With TRACK the function also reports whether the coordinate is free
// (alpha = 1) or pinned at a kink / bound (alpha = 0) and on which side of shift i it sits
// (eps[i] = +-1); the simplex-Newton dual solver builds its generalised Hessian from these.
// Branch-free: it runs inside the four-slot sweep bodies.
template <int KIND, int M, int GF, bool TRACK>
__device__ __forceinline__ double prox_elem(const zf_problem& P, int j, double v,
                                            const double (&wt)[M], double& alpha,
                                            double (&eps)[M], unsigned& code) {
  // code (TRACK): which linear piece of the chain the coordinate is on -- two bits per L1 stage
  // (2: pinned at that stage's kink, 1 / 0: above / below it) and two for the box (1: clipped at
  // the upper bound, 2: at the lower bound).  Two points with equal codes in every coordinate
  // lie on ONE quadratic piece of the dual; pinned at a different kink or bound is a different
  // piece even though alpha is 0 in both.
  constexpr bool L1 = (GF & ZF_G_L1) != 0, BOX = (GF & ZF_G_BOX) != 0;
  double p = v;
  bool pinned = false;
  code = 0u;
#pragma unroll
  for (int i = 0; i < M; ++i) eps[i] = 0.0;
  if constexpr (KIND == ZF_LSQ_L1) {
    double wsum = 0.0;
#pragma unroll
    for (int i = 0; i < M; ++i) wsum += wt[i];
    const double t = P.l1 * wsum;
    p = soft_threshold(v, t);
    if (TRACK) {
      pinned = fabs(v) <= t;
#pragma unroll
      for (int i = 0; i < M; ++i) eps[i] = v > t ? 1.0 : -1.0;
      code = pinned ? 2u : (v > t ? 1u : 0u);
    }
    alpha = pinned ? 0.0 : 1.0;
    return p;
  }
  if constexpr (L1) {
    double coef[M];
#pragma unroll
    for (int i = 0; i < M; ++i) coef[i] = wt[i] * P.l1_ratios[i];
    double tail = 0.0;
#pragma unroll
    for (int i = 1; i < M; ++i) tail += coef[i];
    const double a0 = p + tail - P.l1_shifts[0] + P.l1_shifts[0];
    p = soft_threshold(a0, coef[0]);
    if (TRACK) {
      pinned = fabs(a0) <= coef[0];
      eps[0] = a0 > coef[0] ? 1.0 : -1.0;
      code = pinned ? 2u : (a0 > coef[0] ? 1u : 0u);
    }
#pragma unroll
    for (int i = 1; i < M; ++i) {
      const double ai = p - coef[i] - P.l1_shifts[i];
      p = soft_threshold(ai, coef[i]) + P.l1_shifts[i];
      if (TRACK) {
        const bool pin_i = fabs(ai) <= coef[i];
        pinned = pinned || pin_i;
        eps[i] = ai > coef[i] ? 1.0 : -1.0;
        code |= (pin_i ? 2u : (ai > coef[i] ? 1u : 0u)) << (2 * i);
      }
    }
  }
  if constexpr (BOX) {
    // projection_box (scalar or per-coordinate bounds)
    const double q = fmin(fmax(p, lower_of(P, j)), upper_of(P, j));
    if (TRACK) {
      pinned = pinned || (q != p);
      code |= (q < p ? 1u : (q > p ? 2u : 0u)) << (2 * M);
    }
    p = q;
  }
  alpha = pinned ? 0.0 : 1.0;
  return p;
}

}  // namespace zf
