// Batched FISTA / ISTA: the whole minimize_proximal_gradient loop
// (proximal_gradient.py:474-555) on device, one warp per starting point.
//
// Per outer iteration a warp does, without leaving the SM:
//   f(y), jac_f(y)                     -> registers / shared memory
//   backtracking on lr                 (proximal_gradient.py:279-308)
//     dual solve of the subproblem     (zf_dual.cuh)
//     x = prox(...), F(x) = f(x)+g(x)
//   stopping test  max|x - y| < tol
//   t_{k+1}(a, b), extrapolation y = x + (t_k - 1)/t_{k+1} (x - x_prev)
// State (y, x_prev, x, J rows) lives in the warp's slice of shared memory.
#pragma once
#include <cstdio>
#include <type_traits>

#include "zf_dual.cuh"
#include "zf_host.h"

namespace zf {

__host__ __device__ __forceinline__ size_t warp_smem_doubles(int n, int m, int n_rows) {
  return (size_t)(3 + m) * n + n_rows + (size_t)(n + 3) / 4;     // + n 16-bit piece codes
}

template <int M>
struct SubproblemOut {
  double fun;       // primal subproblem value (= D(w*), res.fun of _solve_subproblem)
  double w[M];
  int n_dual;
};

// What the line search and the stopping test need to know about a candidate x (in c.xn)
template <int M>
struct StepEval {
  double fx[M];     // f(x)
  double gx[M];     // g(x)
  double err;       // max |x - y|
  bool valid;       // filled for the x now in c.xn
};

// f, g and max|x - y| of the x in c.xn by separate passes (single-objective path)
template <int KIND, int M>
__device__ void F_eval_plain(const zf_problem& P, const WarpCtx& c,
                             const typename Fn<KIND, M>::Consts& K, StepEval<M>& ev) {
  f_eval<KIND, M>(P, c, K, c.xn, ev.fx);
  g_eval<KIND, M>(P, c, c.xn, ev.gx);
  double e = 0.0;
#pragma unroll 1
  for (int j = c.lane; j < c.n; j += 32) e = fmax(e, fabs(c.xn[j] - c.y[j]));
  ev.err = warp_max(e);
  ev.valid = true;
}

// x = prox_wsum_g(lr * w, y - lr * w @ J) into c.xn (proximal_gradient.py:206) and, in the SAME
// sweep over the coordinates, everything the outer loop wants to know about that x: the sums
// behind f(x) and g(x) (proximal_gradient.py:283-296) and max|x - y| (:511), with ONE K-sum
// reduction.  With PROBE it also reports whether every coordinate sits on the linear piece of
// the prox chain stored at the last full dual evaluation (zf_dual.cuh:primal_probe).  Every sum
// is accumulated over the same coordinates in the same order as the separate passes
// (primal_probe, f, g_eval, the error loop) did: results are bit-identical to theirs.
template <int KIND, int M, int GF, bool PROBE>
__device__ bool fused_primal_eval(const zf_problem& P, const WarpCtx& c,
                                  const typename Fn<KIND, M>::Consts& K, double lr,
                                  const double (&w)[M], StepEval<M>& ev) {
  using F = Fn<KIND, M>;
  constexpr bool L1 = (GF & ZF_G_L1) != 0, BOX = (GF & ZF_G_BOX) != 0;
  bool same_piece = true;
  if constexpr (!F::kPointwise) {
    if constexpr (PROBE) same_piece = primal_probe<KIND, M, GF>(P, c, lr, w, c.xn);
    else primal_from_weights<KIND, M, GF>(P, c, lr, w, c.xn);
    F_eval_plain<KIND, M>(P, c, K, ev);
  } else {
    double wt[M];
#pragma unroll
    for (int i = 0; i < M; ++i) wt[i] = lr * w[i];
    constexpr int NF = F::NF;
    double sf[NF], sg[M];
    double e;
    int same, bad;
    // the sweep proper; FAST = branch-free exp inside f_pre (see zf_common.cuh), re-run with the
    // library forms if any coordinate was outside the fast path's range
    auto run = [&](auto fast) -> int {
      constexpr bool FAST = decltype(fast)::value;
#pragma unroll
      for (int k = 0; k < NF; ++k) sf[k] = 0.0;
#pragma unroll
      for (int i = 0; i < M; ++i) sg[i] = 0.0;
      e = 0.0;
      same = 1;
      bad = 0;
      int rare = 0;
      struct In { CoordIn<M> c; unsigned short pat; };
      sweep3<In, double>(
          c.n, c.lane,
          [&](int j, bool live) {
            const int jc = live ? j : 0;
            return In{load_coord<M>(c, jc), PROBE ? c.pat[jc] : (unsigned short)0};
          },
          [&](const In& in, int j, bool live) {
            const int jc = live ? j : 0;
            double wj = 0.0;
#pragma unroll
            for (int i = 0; i < M; ++i) wj += w[i] * in.c.J[i];
            const double yj = in.c.y;
            const double v = yj - lr * wj;
            double alpha, eps[M];
            unsigned pcode;
            const double p = prox_elem<KIND, M, GF, PROBE>(P, jc, v, wt, alpha, eps, pcode);
            if constexpr (PROBE) same &= live ? (int)((unsigned short)pcode == in.pat) : 1;
            double t[F::NT];
            F::template f_pre<FAST>(c, jc, p, t, rare);
            F::f_acc(live, t, sf);
            // (p is already clipped to the box, so g's +inf branch can only fire on NaN bounds;
            // the test is kept because Problem.g makes it, problems.py:101-106)
            if constexpr (BOX) {
              const bool out_of_box = (p < lower_of(P, jc)) || (p > upper_of(P, jc));
              bad |= (live && out_of_box) ? 1 : 0;
            }
            if constexpr (L1) {
#pragma unroll
              for (int i = 0; i < M; ++i) sg[i] += msk(live, fabs(p - P.l1_shifts[i]));
            }
            e = fmax(e, msk(live, fabs(p - yj)));
            return p;
          },
          [&](int j, bool live, double p) {
            if (live) c.xn[j] = p;
          });
      return rare;
    };
    if (__any_sync(ZF_FULL_MASK, run(std::true_type{}))) run(std::false_type{});
    __syncwarp();
    if constexpr (F::kFoldable) {
      double s[NF + M];
#pragma unroll
      for (int k = 0; k < NF; ++k) s[k] = sf[k];
#pragma unroll
      for (int i = 0; i < M; ++i) s[NF + i] = sg[i];
      warp_sum_k<NF + M>(s);
#pragma unroll
      for (int k = 0; k < NF; ++k) sf[k] = s[k];
#pragma unroll
      for (int i = 0; i < M; ++i) sg[i] = s[NF + i];
    } else {        // fixed-size problems: f_finish reads x[0..3], only the g sums are reduced
      warp_sum_k<M>(sg);
    }
    ev.err = warp_max(e);
    if constexpr (PROBE) same_piece = __all_sync(ZF_FULL_MASK, same) != 0;
    F::f_finish(P, c, K, c.xn, sf, ev.fx);
    const bool any_bad = BOX && __any_sync(ZF_FULL_MASK, bad);
#pragma unroll
    for (int i = 0; i < M; ++i) {
      double gi = 0.0;
      if (any_bad) gi = CUDART_INF;
      else if (L1) gi = P.l1_ratios[i] * sg[i];
      ev.gx[i] = gi;
    }
  }
  ev.valid = true;
  return same_piece;
}

// _solve_subproblem (proximal_gradient.py:35-209) given f(y), J(y) already in ctx.
// Writes x into c.xn and, with EVAL, its F / error summary into ev.
template <int KIND, int M, int GF, bool EVAL>
__device__ void solve_subproblem(const zf_problem& P, const zf_options& O, const WarpCtx& c,
                                 const typename Fn<KIND, M>::Consts& K, double lr, const double (&fy)[M], const double (&Fprev)[M],
                                 bool deprecated, SubproblemOut<M>& out, StepEval<M>& ev) {
  ev.valid = false;
  if constexpr (M == 1) {
    // x = prox(lr, y - lr * jac); fun = jac.(x - y) + g(x) + ||x - y||^2 / 2 / lr (+ f_y - F_prev)
    double wt[1] = {lr};
    double s[2] = {0.0, 0.0};
#pragma unroll 1
    for (int j = c.lane; j < c.n; j += 32) {
      const double yj = c.y[j];
      const double gj = c.J[j];
      double alpha, eps[1];
            unsigned pcode;
      const double p = prox_elem<KIND, 1, GF, false>(P, j, yj - lr * gj, wt, alpha, eps, pcode);
      c.xn[j] = p;
      s[0] += gj * (p - yj);
      s[1] += (p - yj) * (p - yj);
    }
    __syncwarp();
    warp_sum_k<2>(s);
    double gx[1];
    g_eval<KIND, 1>(P, c, c.xn, gx);
    double fun = s[0] + gx[0] + norm_sq_like_numpy(s[1]) / 2.0 / lr;
    if (!deprecated) fun += fy[0] - Fprev[0];
    out.fun = fun;
    out.w[0] = 1.0;
    out.n_dual = 1;
  } else {
    DualData<M> d;
    d.lr = lr;
    d.use_c = !deprecated;
#pragma unroll
    for (int i = 0; i < M; ++i) d.c[i] = fy[i] - Fprev[i];
    int nf = 0;
    if (M == 2 && O.dual_solver == 0) {
      if constexpr (M == 2) {
        double fmin;
        const double xf = dual_brent<KIND, GF>(P, c, d, O.tol_internal, O.max_iter_internal, &fmin, &nf);
        out.w[0] = xf;
        out.w[1] = 1.0 - xf;
        out.fun = -fmin;
      }
    } else {
      bool x_ready = false;
      out.fun = dual_newton<KIND, M, GF>(P, c, d, out.w, 60, &nf, &x_ready,
                                         [&](const double (&wn)[M]) {
        if constexpr (EVAL) return fused_primal_eval<KIND, M, GF, true>(P, c, K, lr, wn, ev);
        else return primal_probe<KIND, M, GF>(P, c, lr, wn, c.xn);
      });
      if (!x_ready) ev.valid = false;      // c.xn holds a rejected candidate's x
    }
    out.n_dual = nf;
  }
  if (!ev.valid) {
    if constexpr (EVAL) {
      if constexpr (M == 1) {
        F_eval_plain<KIND, M>(P, c, K, ev);
      } else {
        fused_primal_eval<KIND, M, GF, false>(P, c, K, lr, out.w, ev);
      }
    } else if constexpr (M > 1) {
      primal_from_weights<KIND, M, GF>(P, c, lr, out.w, c.xn);
    }
  }
}

template <int KIND, int M, int GF>
__global__ void __launch_bounds__(128, 1)
batched_fista_kernel(zf_problem P, zf_options O, long long n_starts, const double* __restrict__ x0,
                     const double* __restrict__ ab, zf_result R) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  const int n = P.n_features;
  const int n_rows = (KIND == ZF_LSQ_L1) ? P.n_rows : 0;
  double* base = smem + (size_t)warp_in_block * warp_smem_doubles(n, M, n_rows);

  WarpCtx c;
  c.lane = lane;
  c.n = n;
  c.y = base;
  c.xp = base + n;
  c.xn = base + 2 * (size_t)n;
  c.J = base + 3 * (size_t)n;
  c.scratch = base + (size_t)(3 + M) * n;
  c.pat = reinterpret_cast<unsigned short*>(c.scratch + n_rows);

  using F = Fn<KIND, M>;
  const typename F::Consts K = F::make_consts(c);
  const long long total_warps = (long long)gridDim.x * warps_per_block;
  for (long long s = (long long)blockIdx.x * warps_per_block + warp_in_block; s < n_starts;
       s += total_warps) {
    const double* xs = x0 + s * n;
    // return_all traces: dense (trace_capacity entries per start) or ragged (zf_result.trace_offsets)
    const bool ragged = R.trace_offsets != nullptr;
    const long long cap = ragged ? (long long)(R.trace_offsets[s + 1] - R.trace_offsets[s])
                                 : (long long)O.trace_capacity;
    const bool tracing = ragged || cap > 0;
    const long long ebase = ragged ? (long long)R.trace_offsets[s] : s * cap;           // errors
    const long long fbase = ragged ? (long long)R.trace_offsets[s] + s : s * (cap + 1);  // F, x
#pragma unroll 1
    for (int j = lane; j < n; j += 32) {
      const double v = xs[j];
      c.y[j] = v;
      c.xp[j] = v;
      c.xn[j] = v;
      if (tracing && R.allvecs) R.allvecs[fbase * n + j] = v;
    }
    __syncwarp();
    const double na = ab ? ab[2 * s] : O.nesterov_a;
    const double nb = ab ? ab[2 * s + 1] : O.nesterov_b;
    double lr = O.lr;
    double t_prev = 1.0;
    double Fprev[M], Fx[M], fx[M], fy[M], gx[M];
    f_eval<KIND, M>(P, c, K, c.xp, fx);
    g_eval<KIND, M>(P, c, c.xp, gx);
#pragma unroll
    for (int i = 0; i < M; ++i) {
      Fprev[i] = fx[i] + gx[i];
      Fx[i] = Fprev[i];
    }
    if (tracing && R.allfuns && lane == 0) {
#pragma unroll
      for (int i = 0; i < M; ++i) R.allfuns[fbase * M + i] = Fprev[i];
    }
    long long nfev = 1, ndual = 0;
    double wwarm[M];
#pragma unroll
    for (int i = 0; i < M; ++i) wwarm[i] = 1.0 / (double)M;

    int status = 0;          // max_iter reached unless set otherwise
    long long nit = 0;
    double err = CUDART_INF;
    bool failed = false;
    // extrapolation folded into the next f_jac: y = xp + mom (xp - xn) with xp = x^k, xn = x^{k-1}
    // after the swap at the end of an iteration (first iteration: y = x0)
    double mom = 0.0;
    bool extrapolate = false;
    for (long long it = 1; it <= O.max_iter; ++it) {
      nit = it;
      if constexpr (F::kFoldable) {
        F::f_jac(P, c, K, YFold{c.y, c.xp, c.xn, mom, extrapolate}, c.J, fy);
      } else {
        F::f_jac(P, c, K, YPlain{c.y}, c.J, fy);
      }
      __syncwarp();
      ++nfev;
      // ---- backtracking line search ----
      bool found = false;
      SubproblemOut<M> sub;
      StepEval<M> ev;
      for (int bt = 0; bt < O.max_backtrack_iter; ++bt) {
#pragma unroll
        for (int i = 0; i < M; ++i) sub.w[i] = wwarm[i];
        solve_subproblem<KIND, M, GF, true>(P, O, c, K, lr, fy, Fprev, O.deprecated != 0, sub, ev);
        ndual += sub.n_dual;
        ++nfev;
#pragma unroll
        for (int i = 0; i < M; ++i) {
          fx[i] = ev.fx[i];
          Fx[i] = ev.fx[i] + ev.gx[i];
        }
        // The reference passes w0 = 1/m to its inner solver unless warm_start is set.  The
        // simplex Newton solver converges to the same (exact) maximiser from any start, so it
        // always continues from the previous subproblem's weights: near convergence that is
        // one or two dual evaluations instead of three or four.
        if (O.warm_start || !(M == 2 && O.dual_solver == 0)) {
#pragma unroll
          for (int i = 0; i < M; ++i) wwarm[i] = sub.w[i];
        }
        if (O.decay_rate == 1.0) { found = true; break; }
        bool ok = true;
        if (O.deprecated) {
#pragma unroll
          for (int i = 0; i < M; ++i) ok = ok && (fx[i] - fy[i] <= sub.fun + O.tol_internal);
        } else {
#pragma unroll
          for (int i = 0; i < M; ++i) ok = ok && (Fx[i] - Fprev[i] <= sub.fun + O.tol_internal);
        }
        if (ok) { found = true; break; }
        lr *= O.decay_rate;
      }
      if (!found) {
        // RuntimeError("Backtracking failed...") -> x = x_prev, nit - 1 (proximal_gradient.py:493-509)
        failed = true;
        nit = it - 1;
        break;
      }
      err = ev.err;
      if (tracing && it <= cap) {
        if (R.allerrs && lane == 0) R.allerrs[ebase + (it - 1)] = err;
        if (R.allfuns && lane == 0) {
#pragma unroll
          for (int i = 0; i < M; ++i) R.allfuns[(fbase + it) * M + i] = Fx[i];
        }
        if (R.allvecs) {
#pragma unroll 1
          for (int j = lane; j < n; j += 32) R.allvecs[(fbase + it) * n + j] = c.xn[j];
        }
      }
      if (err < O.tol) { status = 1; break; }
      if (it == O.max_iter) break;   // keep x = x^k as the reference's for/else does
      // ---- momentum and extrapolation (proximal_gradient.py:530-538) ----
      if (O.nesterov) {
        const double t_new = sqrt(t_prev * t_prev - na * t_prev + nb) + 0.5;
        mom = (t_prev - 1.0) / t_new;
        t_prev = t_new;
        extrapolate = true;
      }
      if constexpr (!F::kFoldable) {
        if (O.nesterov) {
#pragma unroll 1
          for (int j = lane; j < n; j += 32) {
            const double xj = c.xn[j];
            c.y[j] = xj + mom * (xj - c.xp[j]);
          }
        } else {
          // y = x_prev = x^k
#pragma unroll 1
          for (int j = lane; j < n; j += 32) c.y[j] = c.xn[j];
        }
      }
      double* tmp = c.xp; c.xp = c.xn; c.xn = tmp;
      __syncwarp();
#pragma unroll
      for (int i = 0; i < M; ++i) Fprev[i] = Fx[i];
    }
    // ---- results ----
    const double* xres = failed ? c.xp : c.xn;
    if (failed) {
      status = -1;
#pragma unroll
      for (int i = 0; i < M; ++i) Fx[i] = Fprev[i];
    }
#pragma unroll 1
    for (int j = lane; j < n; j += 32) R.x[s * n + j] = xres[j];
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < M; ++i) R.fun[s * M + i] = Fx[i];
      R.nit[s] = nit;
      R.status[s] = status;
      if (R.lr) R.lr[s] = lr;
      if (R.nfev) R.nfev[s] = nfev;
      if (R.n_dual) R.n_dual[s] = ndual;
      if (R.err) R.err[s] = err;
    }
    __syncwarp();
  }
}

// One subproblem per warp: _solve_subproblem(f, g, jac_f, prox, lr, xk_old, yk, w0)
template <int KIND, int M, int GF>
__global__ void __launch_bounds__(128)
subproblem_kernel(zf_problem P, zf_options O, long long n_items, const double* __restrict__ Y,
                  const double* __restrict__ Xold, const double* __restrict__ LR,
                  const int* __restrict__ dep, double* __restrict__ X, double* __restrict__ FUN,
                  double* __restrict__ W) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  const int n = P.n_features;
  const int n_rows = (KIND == ZF_LSQ_L1) ? P.n_rows : 0;
  double* base = smem + (size_t)warp_in_block * warp_smem_doubles(n, M, n_rows);
  WarpCtx c;
  c.lane = lane; c.n = n;
  c.y = base; c.xp = base + n; c.xn = base + 2 * (size_t)n; c.J = base + 3 * (size_t)n;
  c.scratch = base + (size_t)(3 + M) * n;
  c.pat = reinterpret_cast<unsigned short*>(c.scratch + n_rows);
  using F = Fn<KIND, M>;
  const long long s = (long long)blockIdx.x * warps_per_block + warp_in_block;
  if (s >= n_items) return;
#pragma unroll 1
  for (int j = lane; j < n; j += 32) {
    c.y[j] = Y[s * n + j];
    c.xp[j] = Xold[s * n + j];
  }
  __syncwarp();
  double fy[M], fp[M], gp[M], Fprev[M];
  const typename F::Consts K = F::make_consts(c);
  f_eval<KIND, M>(P, c, K, c.xp, fp);
  g_eval<KIND, M>(P, c, c.xp, gp);
#pragma unroll
  for (int i = 0; i < M; ++i) Fprev[i] = fp[i] + gp[i];
  F::f_jac(P, c, K, YPlain{c.y}, c.J, fy);
  __syncwarp();
  SubproblemOut<M> sub;
  StepEval<M> ev;
#pragma unroll
  for (int i = 0; i < M; ++i) sub.w[i] = 1.0 / (double)M;
  const bool deprecated = dep ? (dep[s] != 0) : (O.deprecated != 0);
  solve_subproblem<KIND, M, GF, false>(P, O, c, K, LR[s], fy, Fprev, deprecated, sub, ev);
#pragma unroll 1
  for (int j = lane; j < n; j += 32) X[s * n + j] = c.xn[j];
  if (lane == 0) {
    FUN[s] = sub.fun;
#pragma unroll
    for (int i = 0; i < M; ++i) W[s * M + i] = sub.w[i];
  }
}

// Problem.f / g / jac_f / prox_wsum_g at a batch of points (one warp per point).
template <int KIND, int M, int GF>
__global__ void __launch_bounds__(128)
problem_eval_kernel(zf_problem P, long long n_items, const double* __restrict__ Xin,
                    const double* __restrict__ Win, double* __restrict__ fo,
                    double* __restrict__ go, double* __restrict__ jo, double* __restrict__ po) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  const int n = P.n_features;
  const int n_rows = (KIND == ZF_LSQ_L1) ? P.n_rows : 0;
  double* base = smem + (size_t)warp_in_block * warp_smem_doubles(n, M, n_rows);
  WarpCtx c;
  c.lane = lane; c.n = n;
  c.y = base; c.xp = base + n; c.xn = base + 2 * (size_t)n; c.J = base + 3 * (size_t)n;
  c.scratch = base + (size_t)(3 + M) * n;
  c.pat = reinterpret_cast<unsigned short*>(c.scratch + n_rows);
  using F = Fn<KIND, M>;
  const long long s = (long long)blockIdx.x * warps_per_block + warp_in_block;
  if (s >= n_items) return;
#pragma unroll 1
  for (int j = lane; j < n; j += 32) c.y[j] = Xin[s * n + j];
  __syncwarp();
  double fy[M], gy[M];
  const typename F::Consts K = F::make_consts(c);
  F::f_jac(P, c, K, YPlain{c.y}, c.J, fy);
  __syncwarp();
  g_eval<KIND, M>(P, c, c.y, gy);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < M; ++i) {
      if (fo) fo[s * M + i] = fy[i];
      if (go) go[s * M + i] = gy[i];
    }
  }
  if (jo) {
#pragma unroll 1
    for (int j = lane; j < n; j += 32) {
#pragma unroll
      for (int i = 0; i < M; ++i) jo[(s * M + i) * n + j] = c.J[i * n + j];
    }
  }
  if (po && Win) {
    double wt[M];
#pragma unroll
    for (int i = 0; i < M; ++i) wt[i] = Win[s * M + i];
#pragma unroll 1
    for (int j = lane; j < n; j += 32) {
      double alpha, eps[M];
            unsigned pcode;
      po[s * n + j] = prox_elem<KIND, M, GF, false>(P, j, c.y[j], wt, alpha, eps, pcode);
    }
  }
}

// ---------------------------------------------------------------------------------------
// host-side dispatch
// ---------------------------------------------------------------------------------------
enum class Op { Solve, Subproblem, Eval };

struct LaunchArgs {
  Op op;
  zf_problem P;
  zf_options O;
  long long n_items;
  const double* a0;   // x0 | Y | X
  const double* a1;   // ab | Xold | W
  const double* a2;   // - | LR | -
  const int* i0;      // - | dep | -
  zf_result R;        // Solve outputs
  double* o0; double* o1; double* o2; double* o3;  // Subproblem: X, FUN, W ; Eval: f, g, jac, prox
  cudaStream_t stream;
};

template <int KIND, int M, int GF>
int launch_t(const LaunchArgs& L) {
  const int n = L.P.n_features;
  const int n_rows = (KIND == ZF_LSQ_L1) ? L.P.n_rows : 0;
  const size_t per_warp = warp_smem_doubles(n, M, n_rows) * sizeof(double);
  const size_t smem_cap = 200 * 1024;
  if (per_warp > smem_cap) {
    return zf_fail(ZF_ERR_UNSUPPORTED,
                   "n_features=%d needs %zu B of shared memory per start (limit %zu); "
                   "use the large-n LASSO path for single-objective problems",
                   n, per_warp, smem_cap);
  }
  // Few starts: one warp per block so the warps spread over all SMs / schedulers.
  int wpb = (L.n_items <= 148LL * 16) ? 1 : 4;
  while (wpb > 1 && per_warp * wpb > smem_cap) wpb >>= 1;
  const size_t smem = per_warp * wpb;
  long long blocks = (L.n_items + wpb - 1) / wpb;
  if (L.op == Op::Solve && blocks > 148LL * 64) blocks = 148LL * 64;  // grid-stride beyond
  if (blocks < 1) blocks = 1;
  cudaError_t e;
  if (L.op == Op::Solve) {
    auto k = batched_fista_kernel<KIND, M, GF>;
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return zf_fail_cuda(e, "cudaFuncSetAttribute");
    k<<<(unsigned)blocks, wpb * 32, smem, L.stream>>>(L.P, L.O, L.n_items, L.a0, L.a1, L.R);
  } else if (L.op == Op::Subproblem) {
    auto k = subproblem_kernel<KIND, M, GF>;
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return zf_fail_cuda(e, "cudaFuncSetAttribute");
    k<<<(unsigned)blocks, wpb * 32, smem, L.stream>>>(L.P, L.O, L.n_items, L.a0, L.a1, L.a2, L.i0,
                                                     L.o0, L.o1, L.o2);
  } else {
    auto k = problem_eval_kernel<KIND, M, GF>;
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return zf_fail_cuda(e, "cudaFuncSetAttribute");
    k<<<(unsigned)blocks, wpb * 32, smem, L.stream>>>(L.P, L.n_items, L.a0, L.a1, L.o0, L.o1, L.o2,
                                                     L.o3);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return zf_fail_cuda(e, "kernel launch");
  zf_count_launch();
  return ZF_OK;
}

}  // namespace zf
