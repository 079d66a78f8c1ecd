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
#include <cstdio>
#include <mutex>

#include "zf_dual.cuh"
#include "zf_host.h"

namespace zf {

__host__ __device__ __forceinline__ size_t warp_smem_doubles(int n, int m, int n_rows) {
  return (size_t)(3 + m) * n + n_rows + (size_t)(n + 7) / 8;     // + n pattern bytes
}

template <int M>
struct SubproblemOut {
  double fun;       // primal subproblem value (= D(w*), res.fun of _solve_subproblem)
  double w[M];
  int n_dual;
};

// _solve_subproblem (proximal_gradient.py:35-209) given f(y), J(y) already in ctx.
// Writes x into c.xn.
template <int KIND, int M>
__device__ void solve_subproblem(const zf_problem& P, const zf_options& O, const WarpCtx& c,
                                 double lr, const double (&fy)[M], const double (&Fprev)[M],
                                 bool deprecated, SubproblemOut<M>& out) {
  if constexpr (M == 1) {
    // x = prox(lr, y - lr * jac); fun = jac.(x - y) + g(x) + ||x - y||^2 / 2 / lr (+ f_y - F_prev)
    double wt[1] = {lr};
    double s[2] = {0.0, 0.0};
#pragma unroll 1
    for (int j = c.lane; j < c.n; j += 32) {
      const double yj = c.y[j];
      const double gj = c.J[j];
      double alpha, eps[1];
      const double p = prox_elem<1, false>(P, j, yj - lr * gj, wt, alpha, eps);
      c.xn[j] = p;
      s[0] += gj * (p - yj);
      s[1] += (p - yj) * (p - yj);
    }
    __syncwarp();
    warp_sum_k<2>(s);
    double gx[1];
    g_eval<1>(P, c, c.xn, gx);
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
        const double xf = dual_brent(P, c, d, O.tol_internal, O.max_iter_internal, &fmin, &nf);
        out.w[0] = xf;
        out.w[1] = 1.0 - xf;
        out.fun = -fmin;
      }
      primal_from_weights<M>(P, c, lr, out.w, c.xn);
    } else {
      bool x_ready = false;
      out.fun = dual_newton<M>(P, c, d, out.w, 60, &nf, &x_ready);
      if (!x_ready) primal_from_weights<M>(P, c, lr, out.w, c.xn);
    }
    out.n_dual = nf;
  }
}

template <int KIND, int M>
__global__ void __launch_bounds__(128)
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
  c.pat = reinterpret_cast<unsigned char*>(c.scratch + n_rows);

  using F = Fn<KIND, M>;
  const long long total_warps = (long long)gridDim.x * warps_per_block;
  for (long long s = (long long)blockIdx.x * warps_per_block + warp_in_block; s < n_starts;
       s += total_warps) {
    const double* xs = x0 + s * n;
    const int cap = O.trace_capacity;
#pragma unroll 1
    for (int j = lane; j < n; j += 32) {
      const double v = xs[j];
      c.y[j] = v;
      c.xp[j] = v;
      c.xn[j] = v;
      if (cap > 0 && R.allvecs) R.allvecs[(s * (cap + 1)) * n + j] = v;
    }
    __syncwarp();
    const double na = ab ? ab[2 * s] : O.nesterov_a;
    const double nb = ab ? ab[2 * s + 1] : O.nesterov_b;
    double lr = O.lr;
    double t_prev = 1.0;
    double Fprev[M], Fx[M], fx[M], fy[M], gx[M];
    F::f(P, c, c.xp, fx);
    g_eval<M>(P, c, c.xp, gx);
#pragma unroll
    for (int i = 0; i < M; ++i) {
      Fprev[i] = fx[i] + gx[i];
      Fx[i] = Fprev[i];
    }
    if (cap > 0 && R.allfuns && lane == 0) {
#pragma unroll
      for (int i = 0; i < M; ++i) R.allfuns[(s * (cap + 1)) * M + i] = Fprev[i];
    }
    long long nfev = 1, ndual = 0;
    double wwarm[M];
#pragma unroll
    for (int i = 0; i < M; ++i) wwarm[i] = 1.0 / (double)M;

    int status = 0;          // max_iter reached unless set otherwise
    long long nit = 0;
    double err = CUDART_INF;
    bool failed = false;
    for (long long it = 1; it <= O.max_iter; ++it) {
      nit = it;
      F::f_jac(P, c, c.y, c.J, fy);
      __syncwarp();
      ++nfev;
      // ---- backtracking line search ----
      bool found = false;
      SubproblemOut<M> sub;
      for (int bt = 0; bt < O.max_backtrack_iter; ++bt) {
#pragma unroll
        for (int i = 0; i < M; ++i) sub.w[i] = wwarm[i];
        solve_subproblem<KIND, M>(P, O, c, lr, fy, Fprev, O.deprecated != 0, sub);
        ndual += sub.n_dual;
        F::f(P, c, c.xn, fx);
        g_eval<M>(P, c, c.xn, gx);
        ++nfev;
#pragma unroll
        for (int i = 0; i < M; ++i) Fx[i] = fx[i] + gx[i];
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
      double e = 0.0;
#pragma unroll 1
      for (int j = lane; j < n; j += 32) e = fmax(e, fabs(c.xn[j] - c.y[j]));
      err = warp_max(e);
      if (cap > 0 && it <= cap) {
        if (R.allerrs && lane == 0) R.allerrs[s * cap + (it - 1)] = err;
        if (R.allfuns && lane == 0) {
#pragma unroll
          for (int i = 0; i < M; ++i) R.allfuns[(s * (cap + 1) + it) * M + i] = Fx[i];
        }
        if (R.allvecs) {
#pragma unroll 1
          for (int j = lane; j < n; j += 32) R.allvecs[(s * (cap + 1) + it) * n + j] = c.xn[j];
        }
      }
      if (err < O.tol) { status = 1; break; }
      if (it == O.max_iter) break;   // keep x = x^k as the reference's for/else does
      // ---- momentum and extrapolation (proximal_gradient.py:530-538) ----
      if (O.nesterov) {
        const double t_new = sqrt(t_prev * t_prev - na * t_prev + nb) + 0.5;
        const double mom = (t_prev - 1.0) / t_new;
#pragma unroll 1
        for (int j = lane; j < n; j += 32) {
          const double xj = c.xn[j];
          c.y[j] = xj + mom * (xj - c.xp[j]);
        }
        t_prev = t_new;
        double* tmp = c.xp; c.xp = c.xn; c.xn = tmp;
      } else {
        // y = x_prev = x^k
#pragma unroll 1
        for (int j = lane; j < n; j += 32) c.y[j] = c.xn[j];
        double* tmp = c.xp; c.xp = c.xn; c.xn = tmp;
      }
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
template <int KIND, int M>
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
  c.pat = reinterpret_cast<unsigned char*>(c.scratch + n_rows);
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
  F::f(P, c, c.xp, fp);
  g_eval<M>(P, c, c.xp, gp);
#pragma unroll
  for (int i = 0; i < M; ++i) Fprev[i] = fp[i] + gp[i];
  F::f_jac(P, c, c.y, c.J, fy);
  __syncwarp();
  SubproblemOut<M> sub;
#pragma unroll
  for (int i = 0; i < M; ++i) sub.w[i] = 1.0 / (double)M;
  const bool deprecated = dep ? (dep[s] != 0) : (O.deprecated != 0);
  solve_subproblem<KIND, M>(P, O, c, LR[s], fy, Fprev, deprecated, sub);
#pragma unroll 1
  for (int j = lane; j < n; j += 32) X[s * n + j] = c.xn[j];
  if (lane == 0) {
    FUN[s] = sub.fun;
#pragma unroll
    for (int i = 0; i < M; ++i) W[s * M + i] = sub.w[i];
  }
}

// Problem.f / g / jac_f / prox_wsum_g at a batch of points (one warp per point).
template <int KIND, int M>
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
  c.pat = reinterpret_cast<unsigned char*>(c.scratch + n_rows);
  using F = Fn<KIND, M>;
  const long long s = (long long)blockIdx.x * warps_per_block + warp_in_block;
  if (s >= n_items) return;
#pragma unroll 1
  for (int j = lane; j < n; j += 32) c.y[j] = Xin[s * n + j];
  __syncwarp();
  double fy[M], gy[M];
  F::f_jac(P, c, c.y, c.J, fy);
  __syncwarp();
  g_eval<M>(P, c, c.y, gy);
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
      po[s * n + j] = prox_elem<M, false>(P, j, c.y[j], wt, alpha, eps);
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

template <int KIND, int M>
static int launch_t(const LaunchArgs& L) {
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
    auto k = batched_fista_kernel<KIND, M>;
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return zf_fail_cuda(e, "cudaFuncSetAttribute");
    k<<<(unsigned)blocks, wpb * 32, smem, L.stream>>>(L.P, L.O, L.n_items, L.a0, L.a1, L.R);
  } else if (L.op == Op::Subproblem) {
    auto k = subproblem_kernel<KIND, M>;
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return zf_fail_cuda(e, "cudaFuncSetAttribute");
    k<<<(unsigned)blocks, wpb * 32, smem, L.stream>>>(L.P, L.O, L.n_items, L.a0, L.a1, L.a2, L.i0,
                                                     L.o0, L.o1, L.o2);
  } else {
    auto k = problem_eval_kernel<KIND, M>;
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

static int validate_problem(const zf_problem& P) {
  if (P.n_features < 1) return zf_fail(ZF_ERR_INVALID, "n_features must be >= 1");
  if (P.n_objectives < 1 || P.n_objectives > ZF_MAX_OBJECTIVES)
    return zf_fail(ZF_ERR_UNSUPPORTED, "n_objectives=%d not supported on device (1..%d)",
                   P.n_objectives, ZF_MAX_OBJECTIVES);
  auto need = [&](int n, int m) -> int {
    if (P.n_features != n || P.n_objectives != m)
      return zf_fail(ZF_ERR_INVALID, "problem kind %d requires n_features=%d n_objectives=%d",
                     P.kind, n, m);
    return ZF_OK;
  };
  switch (P.kind) {
    case ZF_JOS1: if (P.n_objectives != 2) return zf_fail(ZF_ERR_INVALID, "JOS1 has 2 objectives"); break;
    case ZF_SD: return need(4, 2);
    case ZF_FDS: if (P.n_objectives != 3) return zf_fail(ZF_ERR_INVALID, "FDS has 3 objectives"); break;
    case ZF_ZDT1:
      if (P.n_objectives != 2) return zf_fail(ZF_ERR_INVALID, "ZDT1 has 2 objectives");
      if (P.n_features < 2) return zf_fail(ZF_ERR_INVALID, "ZDT1 needs n_features >= 2");
      break;
    case ZF_TOI4: return need(4, 2);
    case ZF_TRIDIA: return need(3, 3);
    case ZF_LFR1: break;
    case ZF_LSQ_L1:
      if (!P.A || !P.b || P.n_rows < 1) return zf_fail(ZF_ERR_INVALID, "LSQ_L1 needs A, b, n_rows");
      if (P.n_objectives > 3) return zf_fail(ZF_ERR_UNSUPPORTED, "LSQ_L1 supports 1..3 objectives");
      break;
    default: return zf_fail(ZF_ERR_INVALID, "unknown problem kind %d", P.kind);
  }
  if (P.has_bounds && P.bounds_are_arrays && (!P.lower_v || !P.upper_v))
    return zf_fail(ZF_ERR_INVALID, "array bounds requested but lower_v/upper_v missing");
  return ZF_OK;
}

int zf_launch(const LaunchArgs& L) {
  int rc = validate_problem(L.P);
  if (rc != ZF_OK) return rc;
  const int m = L.P.n_objectives;
  switch (L.P.kind) {
    case ZF_JOS1: return launch_t<ZF_JOS1, 2>(L);
    case ZF_SD: return launch_t<ZF_SD, 2>(L);
    case ZF_FDS: return launch_t<ZF_FDS, 3>(L);
    case ZF_ZDT1: return launch_t<ZF_ZDT1, 2>(L);
    case ZF_TOI4: return launch_t<ZF_TOI4, 2>(L);
    case ZF_TRIDIA: return launch_t<ZF_TRIDIA, 3>(L);
    case ZF_LFR1:
      if (m == 1) return launch_t<ZF_LFR1, 1>(L);
      if (m == 2) return launch_t<ZF_LFR1, 2>(L);
      if (m == 3) return launch_t<ZF_LFR1, 3>(L);
      return launch_t<ZF_LFR1, 4>(L);
    case ZF_LSQ_L1:
      if (m == 1) return launch_t<ZF_LSQ_L1, 1>(L);
      if (m == 2) return launch_t<ZF_LSQ_L1, 2>(L);
      return launch_t<ZF_LSQ_L1, 3>(L);
  }
  return zf_fail(ZF_ERR_INVALID, "unknown problem kind %d", L.P.kind);
}

}  // namespace zf

// =======================================================================================
// C ABI (include/zfista_b200.h)
// =======================================================================================
using zf::LaunchArgs;
using zf::Op;

static int check_options(const zf_options* o) {
  if (!o) return zf::zf_fail(ZF_ERR_INVALID, "options is NULL");
  if (!(o->lr > 0.0)) return zf::zf_fail(ZF_ERR_INVALID, "lr must be > 0");
  if (o->max_iter < 1) return zf::zf_fail(ZF_ERR_INVALID, "max_iter must be >= 1");
  if (o->max_backtrack_iter < 1) return zf::zf_fail(ZF_ERR_INVALID, "max_backtrack_iter must be >= 1");
  if (!(o->decay_rate > 0.0 && o->decay_rate <= 1.0))
    return zf::zf_fail(ZF_ERR_INVALID, "decay_rate must be in (0, 1]");
  if (o->trace_capacity < 0) return zf::zf_fail(ZF_ERR_INVALID, "trace_capacity must be >= 0");
  return ZF_OK;
}

extern "C" int zf_solve_batched_device(const zf_problem* problem, const zf_options* opt,
                                       int64_t n_starts, const double* d_x0, const double* d_ab,
                                       const zf_result* d_out, void* cuda_stream) {
  if (!problem || !d_x0 || !d_out) return zf::zf_fail(ZF_ERR_INVALID, "NULL argument");
  int rc = check_options(opt);
  if (rc != ZF_OK) return rc;
  if (n_starts < 0) return zf::zf_fail(ZF_ERR_INVALID, "n_starts < 0");
  if (n_starts == 0) return ZF_OK;
  if (!d_out->x || !d_out->fun || !d_out->nit || !d_out->status)
    return zf::zf_fail(ZF_ERR_INVALID, "result.x/fun/nit/status are required");
  LaunchArgs L{};
  L.op = Op::Solve;
  L.P = *problem;
  L.O = *opt;
  L.n_items = n_starts;
  L.a0 = d_x0;
  L.a1 = d_ab;
  L.R = *d_out;
  L.stream = (cudaStream_t)cuda_stream;
  return zf::zf_launch(L);
}

namespace {
#define ZF_CUDA(call)                                               \
  do {                                                              \
    cudaError_t _e = (call);                                        \
    if (_e != cudaSuccess) return zf::zf_fail_cuda(_e, #call);      \
  } while (0)

// Device scratch of the *_host entry points: one grow-only block per process, carved per
// call (cudaMalloc / cudaFree cost ~0.3 ms each and cudaFree synchronises; a 1024-start
// solve is ~1 ms of kernel).  Calls through the host entry points are serialised by the lock.
struct Arena {
  std::mutex mu;
  char* base = nullptr;
  size_t cap = 0, off = 0;
  int device = -1;
  int reset(size_t need) {
    int dev = 0;
    ZF_CUDA(cudaGetDevice(&dev));
    if (dev != device || need > cap) {
      if (base) {
        cudaSetDevice(device);
        cudaFree(base);
        cudaSetDevice(dev);
        base = nullptr;
        cap = 0;
      }
      size_t want = need + (need >> 2) + (1u << 20);
      ZF_CUDA(cudaMalloc((void**)&base, want));
      cap = want;
      device = dev;
    }
    off = 0;
    return ZF_OK;
  }
  void* take(size_t bytes) {
    void* p = base + off;
    off += (bytes + 255) & ~(size_t)255;
    return p;
  }
};
Arena g_arena;
inline size_t pad256(size_t b) { return (b + 255) & ~(size_t)255; }

// a slice of the arena (same interface the RAII buffers had)
struct DevBuf {
  void* p = nullptr;
  void take(size_t bytes) { p = g_arena.take(bytes ? bytes : 8); }
  template <class T> T* as() { return static_cast<T*>(p); }
};

// copies the host-side pointer members of a zf_problem (bounds arrays, A, b) to device
struct DevProblem {
  zf_problem P;
  DevBuf lo, hi, A, b;
  static size_t bytes(const zf_problem& hp) {
    size_t t = 0;
    if (hp.has_bounds && hp.bounds_are_arrays) t += 2 * pad256((size_t)hp.n_features * 8);
    if (hp.kind == ZF_LSQ_L1 && hp.n_rows > 0)
      t += pad256((size_t)hp.n_rows * hp.n_features * 8) + pad256((size_t)hp.n_rows * 8);
    return t + 1024;
  }
  int upload(const zf_problem& hp, cudaStream_t st) {
    P = hp;
    const size_t nb = (size_t)hp.n_features * sizeof(double);
    if (hp.has_bounds && hp.bounds_are_arrays) {
      if (!hp.lower_v || !hp.upper_v) return zf::zf_fail(ZF_ERR_INVALID, "bounds arrays missing");
      lo.take(nb);
      hi.take(nb);
      ZF_CUDA(cudaMemcpyAsync(lo.p, hp.lower_v, nb, cudaMemcpyHostToDevice, st));
      ZF_CUDA(cudaMemcpyAsync(hi.p, hp.upper_v, nb, cudaMemcpyHostToDevice, st));
      P.lower_v = lo.as<double>();
      P.upper_v = hi.as<double>();
    }
    if (hp.kind == ZF_LSQ_L1) {
      if (!hp.A || !hp.b || hp.n_rows < 1) return zf::zf_fail(ZF_ERR_INVALID, "LSQ_L1 needs A, b");
      const size_t ab = (size_t)hp.n_rows * hp.n_features * sizeof(double);
      const size_t bb = (size_t)hp.n_rows * sizeof(double);
      A.take(ab);
      b.take(bb);
      ZF_CUDA(cudaMemcpyAsync(A.p, hp.A, ab, cudaMemcpyHostToDevice, st));
      ZF_CUDA(cudaMemcpyAsync(b.p, hp.b, bb, cudaMemcpyHostToDevice, st));
      P.A = A.as<double>();
      P.b = b.as<double>();
    }
    return ZF_OK;
  }
};
}  // namespace

extern "C" int zf_solve_batched_host(const zf_problem* problem, const zf_options* opt,
                                     int64_t n_starts, const double* h_x0, const double* h_ab,
                                     const zf_result* h_out) {
  if (!problem || !h_x0 || !h_out) return zf::zf_fail(ZF_ERR_INVALID, "NULL argument");
  int rc = check_options(opt);
  if (rc != ZF_OK) return rc;
  if (n_starts < 0) return zf::zf_fail(ZF_ERR_INVALID, "n_starts < 0");
  if (n_starts == 0) return ZF_OK;
  if (!h_out->x || !h_out->fun || !h_out->nit || !h_out->status)
    return zf::zf_fail(ZF_ERR_INVALID, "result.x/fun/nit/status are required");
  rc = zf::zf_require_device();
  if (rc != ZF_OK) return rc;
  cudaStream_t st = 0;
  const size_t N = (size_t)n_starts, n = problem->n_features, m = problem->n_objectives;
  const size_t cap = (size_t)opt->trace_capacity;
  std::lock_guard<std::mutex> lock(g_arena.mu);
  {
    size_t total = DevProblem::bytes(*problem) + 2 * pad256(N * n * 8) + pad256(N * 16) +
                   pad256(N * m * 8) + 6 * pad256(N * 8);
    if (cap > 0)
      total += pad256(N * cap * 8) + pad256(N * (cap + 1) * m * 8) + pad256(N * (cap + 1) * n * 8);
    rc = g_arena.reset(total);
    if (rc != ZF_OK) return rc;
  }
  DevProblem dp;
  rc = dp.upload(*problem, st);
  if (rc != ZF_OK) return rc;
  DevBuf x0, ab, x, fun, nit, status, lr, nfev, ndual, err, allerrs, allfuns, allvecs;
  x0.take(N * n * 8);
  ZF_CUDA(cudaMemcpyAsync(x0.p, h_x0, N * n * 8, cudaMemcpyHostToDevice, st));
  if (h_ab) {
    ab.take(N * 2 * 8);
    ZF_CUDA(cudaMemcpyAsync(ab.p, h_ab, N * 2 * 8, cudaMemcpyHostToDevice, st));
  }
  x.take(N * n * 8);
  fun.take(N * m * 8);
  nit.take(N * 8);
  status.take(N * 4);
  zf_result R{};
  R.x = x.as<double>();
  R.fun = fun.as<double>();
  R.nit = nit.as<int64_t>();
  R.status = status.as<int32_t>();
  if (h_out->lr) { lr.take(N * 8); R.lr = lr.as<double>(); }
  if (h_out->nfev) { nfev.take(N * 8); R.nfev = nfev.as<int64_t>(); }
  if (h_out->n_dual) { ndual.take(N * 8); R.n_dual = ndual.as<int64_t>(); }
  if (h_out->err) { err.take(N * 8); R.err = err.as<double>(); }
  if (cap > 0) {
    if (h_out->allerrs) {
      allerrs.take(N * cap * 8);
      ZF_CUDA(cudaMemsetAsync(allerrs.p, 0, N * cap * 8, st));
      R.allerrs = allerrs.as<double>();
    }
    if (h_out->allfuns) {
      allfuns.take(N * (cap + 1) * m * 8);
      ZF_CUDA(cudaMemsetAsync(allfuns.p, 0, N * (cap + 1) * m * 8, st));
      R.allfuns = allfuns.as<double>();
    }
    if (h_out->allvecs) {
      allvecs.take(N * (cap + 1) * n * 8);
      ZF_CUDA(cudaMemsetAsync(allvecs.p, 0, N * (cap + 1) * n * 8, st));
      R.allvecs = allvecs.as<double>();
    }
  }
  rc = zf_solve_batched_device(&dp.P, opt, n_starts, x0.as<double>(),
                               h_ab ? ab.as<double>() : nullptr, &R, st);
  if (rc != ZF_OK) return rc;
  ZF_CUDA(cudaMemcpyAsync(h_out->x, R.x, N * n * 8, cudaMemcpyDeviceToHost, st));
  ZF_CUDA(cudaMemcpyAsync(h_out->fun, R.fun, N * m * 8, cudaMemcpyDeviceToHost, st));
  ZF_CUDA(cudaMemcpyAsync(h_out->nit, R.nit, N * 8, cudaMemcpyDeviceToHost, st));
  ZF_CUDA(cudaMemcpyAsync(h_out->status, R.status, N * 4, cudaMemcpyDeviceToHost, st));
  if (R.lr) ZF_CUDA(cudaMemcpyAsync(h_out->lr, R.lr, N * 8, cudaMemcpyDeviceToHost, st));
  if (R.nfev) ZF_CUDA(cudaMemcpyAsync(h_out->nfev, R.nfev, N * 8, cudaMemcpyDeviceToHost, st));
  if (R.n_dual) ZF_CUDA(cudaMemcpyAsync(h_out->n_dual, R.n_dual, N * 8, cudaMemcpyDeviceToHost, st));
  if (R.err) ZF_CUDA(cudaMemcpyAsync(h_out->err, R.err, N * 8, cudaMemcpyDeviceToHost, st));
  if (R.allerrs) ZF_CUDA(cudaMemcpyAsync(h_out->allerrs, R.allerrs, N * cap * 8, cudaMemcpyDeviceToHost, st));
  if (R.allfuns) ZF_CUDA(cudaMemcpyAsync(h_out->allfuns, R.allfuns, N * (cap + 1) * m * 8, cudaMemcpyDeviceToHost, st));
  if (R.allvecs) ZF_CUDA(cudaMemcpyAsync(h_out->allvecs, R.allvecs, N * (cap + 1) * n * 8, cudaMemcpyDeviceToHost, st));
  ZF_CUDA(cudaStreamSynchronize(st));
  return ZF_OK;
}

extern "C" int zf_solve_subproblem_host(const zf_problem* problem, const zf_options* opt,
                                        int64_t n, const double* h_y, const double* h_x_old,
                                        const double* h_lr, const int32_t* h_deprecated,
                                        double* h_x, double* h_fun, double* h_weight) {
  if (!problem || !h_y || !h_x_old || !h_lr || !h_x || !h_fun || !h_weight)
    return zf::zf_fail(ZF_ERR_INVALID, "NULL argument");
  int rc = check_options(opt);
  if (rc != ZF_OK) return rc;
  if (n <= 0) return n == 0 ? ZF_OK : zf::zf_fail(ZF_ERR_INVALID, "n < 0");
  rc = zf::zf_require_device();
  if (rc != ZF_OK) return rc;
  cudaStream_t st = 0;
  const size_t N = (size_t)n, nf = problem->n_features, m = problem->n_objectives;
  std::lock_guard<std::mutex> lock(g_arena.mu);
  rc = g_arena.reset(DevProblem::bytes(*problem) + 3 * pad256(N * nf * 8) + 3 * pad256(N * 8) +
                     pad256(N * m * 8));
  if (rc != ZF_OK) return rc;
  DevProblem dp;
  rc = dp.upload(*problem, st);
  if (rc != ZF_OK) return rc;
  DevBuf y, xo, lr, dep, x, fun, w;
  y.take(N * nf * 8);
  xo.take(N * nf * 8);
  lr.take(N * 8);
  x.take(N * nf * 8);
  fun.take(N * 8);
  w.take(N * m * 8);
  ZF_CUDA(cudaMemcpyAsync(y.p, h_y, N * nf * 8, cudaMemcpyHostToDevice, st));
  ZF_CUDA(cudaMemcpyAsync(xo.p, h_x_old, N * nf * 8, cudaMemcpyHostToDevice, st));
  ZF_CUDA(cudaMemcpyAsync(lr.p, h_lr, N * 8, cudaMemcpyHostToDevice, st));
  if (h_deprecated) {
    dep.take(N * 4);
    ZF_CUDA(cudaMemcpyAsync(dep.p, h_deprecated, N * 4, cudaMemcpyHostToDevice, st));
  }
  LaunchArgs L{};
  L.op = Op::Subproblem;
  L.P = dp.P;
  L.O = *opt;
  L.n_items = n;
  L.a0 = y.as<double>();
  L.a1 = xo.as<double>();
  L.a2 = lr.as<double>();
  L.i0 = h_deprecated ? dep.as<int>() : nullptr;
  L.o0 = x.as<double>();
  L.o1 = fun.as<double>();
  L.o2 = w.as<double>();
  L.stream = st;
  rc = zf::zf_launch(L);
  if (rc != ZF_OK) return rc;
  ZF_CUDA(cudaMemcpyAsync(h_x, x.p, N * nf * 8, cudaMemcpyDeviceToHost, st));
  ZF_CUDA(cudaMemcpyAsync(h_fun, fun.p, N * 8, cudaMemcpyDeviceToHost, st));
  ZF_CUDA(cudaMemcpyAsync(h_weight, w.p, N * m * 8, cudaMemcpyDeviceToHost, st));
  ZF_CUDA(cudaStreamSynchronize(st));
  return ZF_OK;
}

extern "C" int zf_problem_eval_host(const zf_problem* problem, int64_t n, const double* h_X,
                                    const double* h_W, double* h_f, double* h_g, double* h_jac,
                                    double* h_prox) {
  if (!problem || !h_X) return zf::zf_fail(ZF_ERR_INVALID, "NULL argument");
  if (h_prox && !h_W) return zf::zf_fail(ZF_ERR_INVALID, "prox output needs weights W");
  if (n <= 0) return n == 0 ? ZF_OK : zf::zf_fail(ZF_ERR_INVALID, "n < 0");
  int rc = zf::zf_require_device();
  if (rc != ZF_OK) return rc;
  cudaStream_t st = 0;
  const size_t N = (size_t)n, nf = problem->n_features, m = problem->n_objectives;
  std::lock_guard<std::mutex> lock(g_arena.mu);
  rc = g_arena.reset(DevProblem::bytes(*problem) + 2 * pad256(N * nf * 8) + 3 * pad256(N * m * 8) +
                     pad256(N * m * nf * 8));
  if (rc != ZF_OK) return rc;
  DevProblem dp;
  rc = dp.upload(*problem, st);
  if (rc != ZF_OK) return rc;
  DevBuf X, W, f, g, jac, prox;
  X.take(N * nf * 8);
  ZF_CUDA(cudaMemcpyAsync(X.p, h_X, N * nf * 8, cudaMemcpyHostToDevice, st));
  if (h_W) {
    W.take(N * m * 8);
    ZF_CUDA(cudaMemcpyAsync(W.p, h_W, N * m * 8, cudaMemcpyHostToDevice, st));
  }
  if (h_f) f.take(N * m * 8);
  if (h_g) g.take(N * m * 8);
  if (h_jac) jac.take(N * m * nf * 8);
  if (h_prox) prox.take(N * nf * 8);
  LaunchArgs L{};
  L.op = Op::Eval;
  L.P = dp.P;
  zf_default_options(&L.O);
  L.n_items = n;
  L.a0 = X.as<double>();
  L.a1 = h_W ? W.as<double>() : nullptr;
  L.o0 = h_f ? f.as<double>() : nullptr;
  L.o1 = h_g ? g.as<double>() : nullptr;
  L.o2 = h_jac ? jac.as<double>() : nullptr;
  L.o3 = h_prox ? prox.as<double>() : nullptr;
  L.stream = st;
  rc = zf::zf_launch(L);
  if (rc != ZF_OK) return rc;
  if (h_f) ZF_CUDA(cudaMemcpyAsync(h_f, f.p, N * m * 8, cudaMemcpyDeviceToHost, st));
  if (h_g) ZF_CUDA(cudaMemcpyAsync(h_g, g.p, N * m * 8, cudaMemcpyDeviceToHost, st));
  if (h_jac) ZF_CUDA(cudaMemcpyAsync(h_jac, jac.p, N * m * nf * 8, cudaMemcpyDeviceToHost, st));
  if (h_prox) ZF_CUDA(cudaMemcpyAsync(h_prox, prox.p, N * nf * 8, cudaMemcpyDeviceToHost, st));
  ZF_CUDA(cudaStreamSynchronize(st));
  return ZF_OK;
}
