// Cameraman-style deblurring (examples/cameraman.ipynb "Objective function" cell):
//     f(x) = || R W x - b ||^2,  g(x) = l1 ||x||_1,  jac_f(x) = 2 W^T R (R W x - b)
// R = correlate2d(., K, mode="same", boundary="symm") with an odd K up to 9x9,
// W = inverse single-level 2-D Haar transform of the coefficient vector [cA,cH,cV,cD].
// The notebook runs minimize_proximal_gradient once per (a, b) momentum pair under joblib;
// here every pair is a RUN of one batched solve.  A run's state machine (line search, stop
// test, t_k, momentum) lives in device memory and is advanced by the LAST CTA of the run to
// finish a round, so a round is
//     fixed step, no F trace :  ONE kernel   (gradient + prox + decision)
//     line search / F trace  :  gradient+prox -> [retry prox] -> F(candidate)+decision
// replayed from a CUDA graph without any host round trip; the host only polls "runs still
// active" once per chunk of rounds.  All reductions have a fixed order: runs are bit
// reproducible.
//
// Tile kernel: CTA = 32x32 image tile of one run.  U = W y on the tile + 2R halo (computed
// from the coefficient vectors, y = x + mom (x - x_prev) formed on the fly), V = R U - b on
// tile + R halo (symmetric reflection at the image border), then R V on the tile and the
// forward Haar transform of each 2x2 block -> gradient coefficients.  Everything between the
// coefficient loads and the gradient store stays in shared memory.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "zf_common.cuh"
#include "zf_host.h"

namespace zf {

constexpr int DB_T = 32;
constexpr int DB_THREADS = 128;
constexpr int DB_STRIP = 8;        // outputs per thread in a vertical strip
constexpr int DB_MAXR = 4;
constexpr int DB_UR = DB_T + 4 * DB_MAXR;   // 48
constexpr int DB_VR = DB_T + 2 * DB_MAXR;   // 40

enum DeblurPhase { DP_INIT = 0, DP_NEW = 1, DP_RETRY = 2, DP_FINAL = 3, DP_DONE = 4 };

struct DeblurRun {
  double lr, t_prev, mom, F_prev, F_x, f_y, sub_fun, err, na, nb;
  long long nit;
  int status, phase, bt, res_buf;
  int cur, prev, nxt, pad;     // rotating roles of the three coefficient buffers
};

struct DeblurDims {
  int H, W, h2, w2, tiles_x, tiles_y, n_tiles, R;
  long long n;        // H * W
  int prox_blocks;
};

struct DeblurSums { double gd, dd, abs1, maxd; };

struct DeblurCtl {
  double tol, tol_internal, decay_rate, l1;
  long long max_iter;
  int max_backtrack, nesterov, deprecated, need_F, cap, finalize, store_yg, pad;
};

struct DeblurBufs {
  double* X[3];
  double *Y, *G;
  const double* b;
  double *fy_part, *fx_part, *abs_part;     // [run][tile]
  DeblurSums* tsum;                         // [run][tile]  prox sums of the gradient kernel
  DeblurSums* psum;                         // [run][block] prox sums of the retry kernel
  DeblurRun* runs;
  unsigned int* tickets;                    // [run]
  double *allerrs, *allfuns;
  unsigned int* n_active;                   // [group]
  int run0, group;                          // this launch covers runs run0 .. run0 + gridDim.y - 1
};

// The blur kernel travels as a kernel PARAMETER (constant bank, indexed with compile-time
// offsets after unrolling): every handle has its own taps, nothing process-wide to race on.
//   k   : the (2R+1)^2 taps, row major (general and column-symmetric forms)
//   a, b: when the kernel is an outer product k[u][v] = a[u] b[v] (to within rounding: every
//         Gaussian-like PSF built as window(...) x window(...), examples/cameraman.ipynb), the
//         SEPARABLE form runs two 1-D passes per correlation: 2 (2R+1) FMAs per pixel, not (2R+1)^2
struct DeblurTaps {
  double k[81];
  double a[9], b[9];
};
enum { DK_GENERAL = 0, DK_SYM = 1, DK_SEP = 2 };

// 1-D horizontal pass of the separable form over `rows` x `cols` outputs:
//   out[r][c] = sum_v tb[v] * in[r][c + v + OFF]
// A thread owns 8 consecutive outputs of one row (16 loads feed 72 FMAs); consecutive lanes own
// consecutive ROWS, so with odd pitches every shared-memory access is conflict free.  `in` and
// `out` may be the same array (the first pass runs in place): every thread first loads all its
// inputs, then the block synchronises, then it stores.  At most two tasks per thread (<= 48 rows
// x 5 strips on 128 threads).
template <int R, int OFF, int IN_PITCH, int OUT_PITCH, bool IN_PLACE>
__device__ __forceinline__ void db_hpass(const double (*in)[IN_PITCH], double (*out)[OUT_PITCH],
                                         int rows, int cols, const double (&tb)[9], int tid) {
  constexpr int K = 2 * R + 1, S = 8, NIN = S + 2 * R;
  const int strips = (cols + S - 1) / S;
  const int n_tasks = rows * strips;
  double reg[2][NIN];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int task = tid + t * DB_THREADS;
    if (task < n_tasks) {
      const int r = task % rows, c0 = S * (task / rows);
#pragma unroll
      for (int k = 0; k < NIN; ++k) reg[t][k] = in[r][c0 + k + OFF];
    }
  }
  if (IN_PLACE) __syncthreads();
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int task = tid + t * DB_THREADS;
    if (task < n_tasks) {
      const int r = task % rows, c0 = S * (task / rows);
      double acc[S];
#pragma unroll
      for (int o = 0; o < S; ++o) acc[o] = 0.0;
#pragma unroll
      for (int v = 0; v < K; ++v) {
        const double w = tb[v];
#pragma unroll
        for (int o = 0; o < S; ++o) acc[o] = fma(w, reg[t][o + v], acc[o]);
      }
#pragma unroll
      for (int o = 0; o < S; ++o)
        if (c0 + o < cols) out[r][c0 + o] = acc[o];
    }
  }
}

__device__ __forceinline__ int db_reflect(int i, int n) {
  int r = i < 0 ? -i - 1 : (i >= n ? 2 * n - i - 1 : i);
  return r < 0 ? 0 : (r >= n ? n - 1 : r);
}

// y = x + mom * (x - x_prev) with numpy's rounding (mul and add separately; identical at
// every site that forms y)
__device__ __forceinline__ double db_extrap(double x, double xp, double mom) {
  return __dadd_rn(x, __dmul_rn(mom, __dsub_rn(x, xp)));
}

// One warp: fixed-order reduction of the round's partials, then the reference's scalar logic
// (proximal_gradient.py:149-155, 279-308, 510-538) by lane 0.
__device__ void deblur_decide(const DeblurDims& d, const DeblurCtl& c, const DeblurBufs& B, int run,
                              int lane, bool count) {
  DeblurRun st = B.runs[run];
  if (st.phase == DP_DONE) return;
  if (st.phase == DP_FINAL && !c.finalize) return;   // waits for the host's final F pass
  double fy = 0.0, fx = 0.0, ax = 0.0;
  for (int t = lane; t < d.n_tiles; t += 32) {
    fy += B.fy_part[(long long)run * d.n_tiles + t];
    fx += B.fx_part[(long long)run * d.n_tiles + t];
    ax += B.abs_part[(long long)run * d.n_tiles + t];
  }
  fy = warp_sum(fy); fx = warp_sum(fx); ax = warp_sum(ax);
  DeblurSums s{0.0, 0.0, 0.0, 0.0};
  const bool from_tiles = (st.phase == DP_NEW);
  const int n_part = from_tiles ? d.n_tiles : d.prox_blocks;
  const DeblurSums* part = from_tiles ? B.tsum + (long long)run * d.n_tiles
                                      : B.psum + (long long)run * d.prox_blocks;
  for (int t = lane; t < n_part; t += 32) {
    const DeblurSums u = part[t];
    s.gd += u.gd; s.dd += u.dd; s.abs1 += u.abs1; s.maxd = fmax(s.maxd, u.maxd);
  }
  s.gd = warp_sum(s.gd); s.dd = warp_sum(s.dd); s.abs1 = warp_sum(s.abs1);
  s.maxd = warp_max(s.maxd);
  if (lane != 0) return;
  // f = np.linalg.norm(.)**2 : sqrt, then square
  const double fyv = norm_sq_like_numpy(fy), fxv = norm_sq_like_numpy(fx);
  bool accepted = false;
  if (st.phase == DP_INIT) {
    st.F_prev = fxv + c.l1 * ax;
    st.F_x = st.F_prev;
    if (c.cap > 0 && B.allfuns) B.allfuns[(long long)run * (c.cap + 1)] = st.F_prev;
    st.nit = 1;
    st.phase = DP_NEW;
  } else if (st.phase == DP_FINAL) {
    st.F_x = fxv + c.l1 * ax;
    st.phase = DP_DONE;
  } else {
    if (st.phase == DP_NEW) { st.f_y = fyv; st.bt = 0; }
    double fun = s.gd + c.l1 * s.abs1 + norm_sq_like_numpy(s.dd) / 2.0 / st.lr;
    if (!c.deprecated) fun += st.f_y - st.F_prev;
    st.sub_fun = fun;
    if (c.need_F) {
      const double Fx = fxv + c.l1 * s.abs1;
      bool ok;
      if (c.decay_rate == 1.0) ok = true;
      else if (c.deprecated) ok = (fxv - st.f_y <= fun + c.tol_internal);
      else ok = (Fx - st.F_prev <= fun + c.tol_internal);
      if (ok) { st.F_x = Fx; accepted = true; }
      else {
        st.lr *= c.decay_rate;
        st.bt += 1;
        if (st.bt >= c.max_backtrack) {          // RuntimeError path: x = x_prev, nit - 1
          st.status = -1; st.res_buf = st.cur; st.F_x = st.F_prev; st.nit -= 1;
          st.phase = DP_DONE;
        } else st.phase = DP_RETRY;
      }
    } else accepted = true;
  }
  if (accepted) {
    st.err = s.maxd;
    if (c.cap > 0 && st.nit <= c.cap) {
      if (B.allerrs) B.allerrs[(long long)run * c.cap + (st.nit - 1)] = st.err;
      if (B.allfuns && c.need_F) B.allfuns[(long long)run * (c.cap + 1) + st.nit] = st.F_x;
    }
    const bool conv = st.err < c.tol;
    if (conv || st.nit >= c.max_iter) {
      st.status = conv ? 1 : 0;
      st.res_buf = st.nxt;
      st.phase = c.need_F ? DP_DONE : DP_FINAL;
    } else {
      double mom = 0.0;
      if (c.nesterov) {
        const double t = st.t_prev;
        const double t_new = sqrt(t * t - st.na * t + st.nb) + 0.5;
        mom = (t - 1.0) / t_new;
        st.t_prev = t_new;
      }
      st.mom = mom;
      // the candidate becomes x^k, x^k becomes x^{k-1}, the old x^{k-1} is the next candidate
      const int old_prev = st.prev;
      st.prev = st.cur; st.cur = st.nxt; st.nxt = old_prev;
      st.F_prev = st.F_x;
      st.nit += 1;
      st.phase = DP_NEW;
    }
  }
  B.runs[run] = st;
  if (count && st.phase != DP_DONE && st.phase != DP_FINAL) atomicAdd(B.n_active + B.group, 1u);
}

// MODE 0: gradient round.  U = W y on tile + 2R halo, V = R U - b on tile + R halo, R V on the
//         tile, gradient coefficients per 2x2 block, and immediately the prox candidate
//         x = soft(y - lr g, lr l1) of those coefficients into the run's spare buffer, with
//         the four sums of the subproblem value.
// MODE 1: F evaluation of a buffer (U halo R, V tile only): f partial + ||x||_1 partial.
// The last CTA of a run to finish (ticket counter) advances the run's state machine when
// `decide` is set.
// SYM: the blur kernel is mirror symmetric in its columns (K[u][v] == K[u][2R-v], true for
// every Gaussian-like PSF): the two mirrored input columns are added first and share their
// multiplications -- 8+2R adds + 8K FMAs per column pair instead of 16K FMAs (-35 % FP64 work).
template <int R, int MODE, int KIND>
__global__ void __launch_bounds__(DB_THREADS, KIND == DK_SEP ? 7 : 1)
deblur_tile_kernel(DeblurDims d, DeblurCtl c, DeblurBufs B, const __grid_constant__ DeblurTaps taps,
                   int decide, int count) {
  constexpr bool SYM = (KIND == DK_SYM);
  constexpr bool SEP = (KIND == DK_SEP);
  constexpr int T = DB_T;
  constexpr int HV = (MODE == 0) ? R : 0;         // halo of V
  // halo of U, rounded up to even so that the region is aligned to the 2x2 Haar blocks
  constexpr int HU = (((MODE == 0) ? 2 * R : R) + 1) & ~1;
  constexpr int OFF = HU - HV - R;                // U index of tap (0,0) of V position (0,0)
  constexpr int UR = T + 2 * HU, VR = T + 2 * HV;
  constexpr int K = 2 * R + 1;
  const int run = B.run0 + blockIdx.y;
  const DeblurRun st = B.runs[run];
  const double *xa, *xb;
  double mom = 0.0;
  if (MODE == 0) {
    if (st.phase != DP_NEW) return;
    xa = B.X[st.cur] + (long long)run * d.n;
    xb = B.X[st.prev] + (long long)run * d.n;
    mom = st.mom;
  } else {
    int buf;
    if (st.phase == DP_INIT) buf = st.cur;
    else if (st.phase == DP_NEW || st.phase == DP_RETRY) buf = st.nxt;
    else if (st.phase == DP_FINAL && c.finalize) buf = st.res_buf;
    else return;
    xa = B.X[buf] + (long long)run * d.n;
    xb = xa;
  }
  // 48 physical rows cover every strip overhang: the last strip of the V region reads rows up
  // to S*(VS-1) + S + 2R - 1 + OFF <= 47 for every R <= 4 and both modes
  __shared__ double U[DB_UR][DB_UR + 1];
  __shared__ double V[DB_VR][DB_VR + 1];
  __shared__ double red[6][DB_THREADS / 32];
  const int tid = threadIdx.x;
  const int tyo = (blockIdx.x / d.tiles_x) * T, txo = (blockIdx.x % d.tiles_x) * T;
  const long long q = (long long)d.h2 * d.w2;
  double abs_acc = 0.0;
  // ---- 0. (separable form) the observed image on the V region goes into V's storage now: these
  // loads fly while the coefficient gather below waits for its own, and step 2 subtracts from
  // shared memory instead of stalling on global loads between its two passes
  // (cp.async: global -> shared without a register in between, so nothing here waits for a
  // load; an out-of-image position copies 0 source bytes, i.e. is zero-filled.  As plain
  // load-then-store this loop stalled once per trip: 15 % of the kernel's stall samples.)
  if constexpr (SEP) {
    for (int idx = tid; idx < VR * VR; idx += DB_THREADS) {
      const int li = idx / VR, lj = idx % VR;
      const int gi = tyo - HV + li, gj = txo - HV + lj;
      const bool in = (gi >= 0 && gi < d.H && gj >= 0 && gj < d.W);
      const double* src = B.b + (in ? (long long)gi * d.W + gj : 0);
      const unsigned dst = (unsigned)__cvta_generic_to_shared(&V[li][lj]);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src),
                   "r"(in ? 8 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  // ---- 1. U = W y on the halo region, by 2x2 blocks (one coefficient quadruple each)
  constexpr int UB = UR / 2;
  // fully unrolled (<= 5 trips): all of a thread's coefficient loads are in flight at once --
  // this gather from L2 is the latency the round-1 profile showed the kernel waiting on
#pragma unroll
  for (int blk = tid; blk < UB * UB; blk += DB_THREADS) {
    const int lbi = blk / UB, lbj = blk % UB;
    const int i0 = tyo - HU + 2 * lbi, j0 = txo - HU + 2 * lbj;
    const int gi0 = db_reflect(i0, d.H), gi1 = db_reflect(i0 + 1, d.H);
    const int gj0 = db_reflect(j0, d.W), gj1 = db_reflect(j0 + 1, d.W);
    const long long o = (long long)(gi0 >> 1) * d.w2 + (gj0 >> 1);
    double ca = xa[o], ch = xa[q + o], cv = xa[2 * q + o], cd = xa[3 * q + o];
    if (MODE == 0) {
      ca = db_extrap(ca, xb[o], mom);
      ch = db_extrap(ch, xb[q + o], mom);
      cv = db_extrap(cv, xb[2 * q + o], mom);
      cd = db_extrap(cd, xb[3 * q + o], mom);
    } else {
      const bool interior = (i0 >= tyo) && (i0 < tyo + T) && (j0 >= txo) && (j0 < txo + T) &&
                            (i0 < d.H) && (j0 < d.W);
      if (interior) abs_acc += fabs(ca) + fabs(ch) + fabs(cv) + fabs(cd);
    }
#pragma unroll
    for (int di = 0; di < 2; ++di) {
      const int gi = di ? gi1 : gi0;
      const double sh = (gi & 1) ? -1.0 : 1.0;
#pragma unroll
      for (int dj = 0; dj < 2; ++dj) {
        const int gj = dj ? gj1 : gj0;
        const double sv = (gj & 1) ? -1.0 : 1.0;
        // idwt2 (haar): (cA +- cH +- cV +- cD) / 2, summed left to right
        U[2 * lbi + di][2 * lbj + dj] = (((ca + sh * ch) + sv * cv) + (sh * sv) * cd) / 2.0;
      }
    }
  }
  if constexpr (SEP) asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // ---- 2. V = R U - b at the in-image positions of the V region.
  // A thread owns DB_STRIP vertically adjacent outputs: per kernel column it loads
  // DB_STRIP + 2R values of U once and reuses them for all K kernel rows (8 x K DFMA per
  // 8 + 2R shared-memory loads; neighbouring lanes read neighbouring columns, conflict free).
  double fsum = 0.0;
  constexpr int S = DB_STRIP;
  constexpr int VS = (VR + S - 1) / S;
  if constexpr (SEP) {
    // horizontal pass in place: U[r][lj] <- sum_v b[v] U[r][lj + v + OFF] on the rows the
    // vertical pass reads (OFF .. OFF + VR + 2R - 1; row indices are shifted by -OFF on the way)
    db_hpass<R, OFF, DB_UR + 1, DB_UR + 1, true>(U + OFF, U + OFF, VR + 2 * R, VR, taps.b, tid);
    __syncthreads();
  }
  for (int task = tid; task < VS * VR; task += DB_THREADS) {
    const int li0 = S * (task / VR), lj = task % VR;
    const int gj = txo - HV + lj;
    const bool col_ok = (gj >= 0 && gj < d.W);
    double acc[S], bv[S];
#pragma unroll
    for (int o = 0; o < S; ++o) {       // observed image values: loaded before the FMA loop
      const int gi = tyo - HV + li0 + o;
      acc[o] = 0.0;
      if constexpr (SEP) bv[o] = (li0 + o < VR) ? V[li0 + o][lj] : 0.0;
      else bv[o] = (col_ok && li0 + o < VR && gi >= 0 && gi < d.H) ? B.b[(long long)gi * d.W + gj] : 0.0;
    }
    if constexpr (SEP) {
      // vertical pass of the separable form: one column of the row-filtered U
      double col[S + 2 * R];
#pragma unroll
      for (int k = 0; k < S + 2 * R; ++k) col[k] = U[li0 + k + OFF][lj];
#pragma unroll
      for (int u = 0; u < K; ++u) {
        const double w = taps.a[u];
#pragma unroll
        for (int o = 0; o < S; ++o) acc[o] = fma(w, col[o + u], acc[o]);
      }
    } else {
#pragma unroll
    for (int v = 0; v < (SYM ? R + 1 : K); ++v) {
      double col[S + 2 * R];
#pragma unroll
      for (int k = 0; k < S + 2 * R; ++k) {
        col[k] = U[li0 + k + OFF][lj + v + OFF];
        if (SYM && v < R) col[k] += U[li0 + k + OFF][lj + (2 * R - v) + OFF];
      }
#pragma unroll
      for (int u = 0; u < K; ++u) {
        const double w = taps.k[u * K + v];
#pragma unroll
        for (int o = 0; o < S; ++o) acc[o] += w * col[o + u];
      }
    }
    }
    if (col_ok) {
#pragma unroll
      for (int o = 0; o < S; ++o) {
        const int li = li0 + o, gi = tyo - HV + li;
        if (li < VR && gi >= 0 && gi < d.H) {
          const double val = acc[o] - bv[o];
          V[li][lj] = val;
          if (li >= HV && li < HV + T && lj >= HV && lj < HV + T) fsum += val * val;
        }
      }
    }
  }
  DeblurSums ps{0.0, 0.0, 0.0, 0.0};
  if (MODE == 0) {
    __syncthreads();
    // ---- 3. symmetric reflection of V into the out-of-image halo positions (border tiles only)
    const bool touches_border = (tyo - HV < 0) || (txo - HV < 0) || (tyo + T + HV > d.H) ||
                                (txo + T + HV > d.W);
    if (touches_border) {
      // only the out-of-image strips are visited: the rows above / below the image over their
      // whole width, then the columns left / right of it over the in-image rows
      const int r_lo = (tyo - HV < 0) ? -(tyo - HV) : 0;
      const int r_hi = (d.H - (tyo - HV) < VR) ? d.H - (tyo - HV) : VR;
      const int c_lo = (txo - HV < 0) ? -(txo - HV) : 0;
      const int c_hi = (d.W - (txo - HV) < VR) ? d.W - (txo - HV) : VR;
      auto fill = [&](int li, int lj) {
        const int gi = tyo - HV + li, gj = txo - HV + lj;
        const int si = db_reflect(gi, d.H) - (tyo - HV), sj = db_reflect(gj, d.W) - (txo - HV);
        if (si >= 0 && si < VR && sj >= 0 && sj < VR) V[li][lj] = V[si][sj];
        else V[li][lj] = 0.0;   // never read by an in-image output
      };
      const int out_rows = r_lo + (VR - r_hi);
      for (int idx = tid; idx < out_rows * VR; idx += DB_THREADS) {
        const int rr = idx / VR, lj = idx % VR;
        fill(rr < r_lo ? rr : r_hi + (rr - r_lo), lj);
      }
      const int out_cols = c_lo + (VR - c_hi);
      for (int idx = tid; idx < (r_hi - r_lo) * out_cols; idx += DB_THREADS) {
        const int rr = idx / out_cols, cc = idx % out_cols;
        fill(r_lo + rr, cc < c_lo ? cc : c_hi + (cc - c_lo));
      }
    }
    __syncthreads();
    // ---- 4. Wimg = R V on the tile (vertical strips of DB_STRIP), into U's storage
    // (separable form: rows of V filtered into U's storage, then the columns of that back into
    // V's storage -- V is dead once its rows have been read -- and step 5 reads Wimg from V)
    if constexpr (SEP) {
      db_hpass<R, 0, DB_VR + 1, DB_UR + 1, false>(V, U, T + 2 * R, T, taps.b, tid);
      __syncthreads();
      for (int task = tid; task < (T / S) * T; task += DB_THREADS) {
        const int li0 = S * (task / T), lj = task % T;
        double acc[S], col[S + 2 * R];
#pragma unroll
        for (int o = 0; o < S; ++o) acc[o] = 0.0;
#pragma unroll
        for (int k = 0; k < S + 2 * R; ++k) col[k] = U[li0 + k][lj];
#pragma unroll
        for (int u = 0; u < K; ++u) {
          const double w = taps.a[u];
#pragma unroll
          for (int o = 0; o < S; ++o) acc[o] = fma(w, col[o + u], acc[o]);
        }
#pragma unroll
        for (int o = 0; o < S; ++o) V[li0 + o][lj] = acc[o];
      }
    } else
    for (int task = tid; task < (T / S) * T; task += DB_THREADS) {
      const int li0 = S * (task / T), lj = task % T;
      double acc[S];
#pragma unroll
      for (int o = 0; o < S; ++o) acc[o] = 0.0;
#pragma unroll
      for (int v = 0; v < (SYM ? R + 1 : K); ++v) {
        double col[S + 2 * R];
#pragma unroll
        for (int k = 0; k < S + 2 * R; ++k) {
          col[k] = V[li0 + k][lj + v];
          if (SYM && v < R) col[k] += V[li0 + k][lj + (2 * R - v)];
        }
#pragma unroll
        for (int u = 0; u < K; ++u) {
          const double w = taps.k[u * K + v];
#pragma unroll
          for (int o = 0; o < S; ++o) acc[o] += w * col[o + u];
        }
      }
#pragma unroll
      for (int o = 0; o < S; ++o) U[li0 + o][lj] = acc[o];
    }
    __syncthreads();
    // ---- 5. gradient coefficients = 2 * dwt2(Wimg) per 2x2 block, and the prox candidate
    double* xn = B.X[st.nxt] + (long long)run * d.n;
    const double lr = st.lr, thr = __dmul_rn(c.l1, lr);
    for (int blk = tid; blk < (T / 2) * (T / 2); blk += DB_THREADS) {
      const int bi = blk / (T / 2), bj = blk % (T / 2);
      const int gi = tyo + 2 * bi, gj = txo + 2 * bj;
      if (gi < d.H && gj < d.W) {
        double p00, p01, p10, p11;
        if constexpr (SEP) {
          p00 = V[2 * bi][2 * bj]; p01 = V[2 * bi][2 * bj + 1];
          p10 = V[2 * bi + 1][2 * bj]; p11 = V[2 * bi + 1][2 * bj + 1];
        } else {
          p00 = U[2 * bi][2 * bj]; p01 = U[2 * bi][2 * bj + 1];
          p10 = U[2 * bi + 1][2 * bj]; p11 = U[2 * bi + 1][2 * bj + 1];
        }
        const long long o = (long long)(gi >> 1) * d.w2 + (gj >> 1);
        double g4[4];
        g4[0] = 2.0 * ((((p00 + p01) + p10) + p11) / 2.0);
        g4[1] = 2.0 * ((((p00 + p01) - p10) - p11) / 2.0);
        g4[2] = 2.0 * ((((p00 - p01) + p10) - p11) / 2.0);
        g4[3] = 2.0 * ((((p00 - p01) - p10) + p11) / 2.0);
#pragma unroll
        for (int sb = 0; sb < 4; ++sb) {
          const long long oo = sb * q + o;
          const double yj = db_extrap(xa[oo], xb[oo], mom);
          const double gj4 = g4[sb];
          const double xj = soft_threshold(fma(-lr, gj4, yj), thr);
          const double dd = xj - yj;
          xn[oo] = xj;
          if (c.store_yg) {
            B.Y[(long long)run * d.n + oo] = yj;
            B.G[(long long)run * d.n + oo] = gj4;
          }
          ps.gd += gj4 * dd;
          ps.dd += dd * dd;
          ps.abs1 += fabs(xj);
          ps.maxd = fmax(ps.maxd, fabs(dd));
        }
      }
    }
  }
  // ---- block reduction of the partials (fixed order), ticket, decision by the last CTA
  {
    // five sums in one recursive-halving reduction (zf_common.cuh: 13 double shuffles instead of
    // 25, bit-identical to five butterflies)
    double v5[5] = {fsum, abs_acc, ps.gd, ps.dd, ps.abs1};
    warp_sum_k<5>(v5);
    fsum = v5[0]; abs_acc = v5[1]; ps.gd = v5[2]; ps.dd = v5[3]; ps.abs1 = v5[4];
  }
  ps.maxd = warp_max(ps.maxd);
  if ((tid & 31) == 0) {
    const int w = tid >> 5;
    red[0][w] = fsum; red[1][w] = abs_acc;
    red[2][w] = ps.gd; red[3][w] = ps.dd; red[4][w] = ps.abs1; red[5][w] = ps.maxd;
  }
  __syncthreads();
  if (tid == 0) {
    double t[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int w = 0; w < DB_THREADS / 32; ++w) {
#pragma unroll
      for (int k = 0; k < 5; ++k) t[k] += red[k][w];
      t[5] = fmax(t[5], red[5][w]);
    }
    const long long slot = (long long)run * d.n_tiles + blockIdx.x;
    if (MODE == 0) {
      B.fy_part[slot] = t[0];
      B.tsum[slot] = DeblurSums{t[2], t[3], t[4], t[5]};
    } else {
      B.fx_part[slot] = t[0];
      B.abs_part[slot] = t[1];
    }
  }
  // ticket + decision: only warp 0 stays for it, the other warps are done
  if (decide && tid < 32) {
    int last = 0;
    if (tid == 0) {
      __threadfence();
      const unsigned int done = atomicAdd(&B.tickets[run], 1u);
      last = (done == (unsigned int)d.n_tiles - 1u);
      if (last) B.tickets[run] = 0u;
    }
    last = __shfl_sync(ZF_FULL_MASK, last, 0);
    if (last) {
      __threadfence();
      deblur_decide(d, c, B, run, tid, count != 0);
    }
  }
}

// Line-search retry: the same gradient, a smaller step.  x = soft(y - lr g, lr l1) from the
// stored y and g into the run's spare buffer + block partials.
__global__ void __launch_bounds__(DB_THREADS)
deblur_prox_kernel(DeblurDims d, DeblurCtl c, DeblurBufs B) {
  const int run = B.run0 + blockIdx.y;
  const DeblurRun st = B.runs[run];
  if (st.phase != DP_RETRY) return;
  double* xn = B.X[st.nxt] + (long long)run * d.n;
  const double* y = B.Y + (long long)run * d.n;
  const double* g = B.G + (long long)run * d.n;
  const double lr = st.lr, thr = __dmul_rn(c.l1, lr);
  DeblurSums s{0.0, 0.0, 0.0, 0.0};
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < d.n;
       j += (long long)gridDim.x * blockDim.x) {
    const double gj = g[j], yj = y[j];
    const double xj = soft_threshold(fma(-lr, gj, yj), thr);
    const double dd = xj - yj;
    xn[j] = xj;
    s.gd += gj * dd;
    s.dd += dd * dd;
    s.abs1 += fabs(xj);
    s.maxd = fmax(s.maxd, fabs(dd));
  }
  __shared__ DeblurSums sh[DB_THREADS / 32];
  s.gd = warp_sum(s.gd);
  s.dd = warp_sum(s.dd);
  s.abs1 = warp_sum(s.abs1);
  s.maxd = warp_max(s.maxd);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    DeblurSums t = sh[0];
    for (int w = 1; w < DB_THREADS / 32; ++w) {
      t.gd += sh[w].gd; t.dd += sh[w].dd; t.abs1 += sh[w].abs1; t.maxd = fmax(t.maxd, sh[w].maxd);
    }
    B.psum[(long long)run * d.prox_blocks + blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(DB_THREADS)
deblur_gather_kernel(DeblurDims d, DeblurBufs B, double* __restrict__ out) {
  const int run = B.run0 + blockIdx.y;
  const double* src = B.X[B.runs[run].res_buf] + (long long)run * d.n;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < d.n;
       j += (long long)gridDim.x * blockDim.x)
    out[(long long)run * d.n + j] = src[j];
}

}  // namespace zf

// =======================================================================================
struct zf_deblur {
  zf::DeblurDims d{};
  zf::DeblurBufs B{};
  double l1 = 0.0;
  int max_runs = 0;
  unsigned int* h_active = nullptr;   // pinned, [group]
  cudaStream_t st2[3] = {nullptr, nullptr, nullptr};   // further groups of runs (see deblur_run)
  cudaEvent_t ev = nullptr;
  double* scratch = nullptr;          // max_runs x n: gathered results
  size_t trace_cap_alloc = 0;
  cudaStream_t st = nullptr;
  cudaStream_t own_st = nullptr;     // created when the caller passes no stream (graphs cannot
                                     // be captured on the legacy default stream)
  bool capturing = false;
  int kind = zf::DK_GENERAL;         // DK_SYM: kernel columns mirror symmetric (exactly): folded
                                     // stencil; DK_SEP: outer product (to rounding): two 1-D passes
  zf::DeblurTaps taps{};
  int device = 0;                    // the device the handle's buffers live on
  std::mutex mu;
};

namespace {
#define ZF_CUDA(call)                                          \
  do {                                                         \
    cudaError_t _e = (call);                                   \
    if (_e != cudaSuccess) return zf::zf_fail_cuda(_e, #call); \
  } while (0)

struct RunGroup {
  cudaStream_t st;
  int run0, n_runs, group;
};

template <int MODE>
int launch_tile_g(zf_deblur* h, const RunGroup& g, const zf::DeblurCtl& c, bool decide, bool count) {
  dim3 grid((unsigned)h->d.n_tiles, (unsigned)g.n_runs);
  const int dd = decide ? 1 : 0, cc = count ? 1 : 0;
  zf::DeblurBufs B = h->B;
  B.run0 = g.run0;
  B.group = g.group;
  // (the separable form is built for 7 CTAs per SM -- 72 registers, 7 x 32.4 KB of shared memory:
  // ask for the largest shared-memory carve-out so that all of them fit)
#define ZF_DB_LAUNCH(RR, KK)                                                                  \
  do {                                                                                        \
    if (KK == zf::DK_SEP)                                                                     \
      cudaFuncSetAttribute(zf::deblur_tile_kernel<RR, MODE, KK>,                               \
                           cudaFuncAttributePreferredSharedMemoryCarveout, 100);              \
    zf::deblur_tile_kernel<RR, MODE, KK><<<grid, zf::DB_THREADS, 0, g.st>>>(h->d, c, B, h->taps, \
                                                                           dd, cc);           \
  } while (0)
#define ZF_DB_LAUNCH_R(KK)                                                                    \
  switch (h->d.R) {                                                                          \
    case 1: ZF_DB_LAUNCH(1, KK); break;                                                       \
    case 2: ZF_DB_LAUNCH(2, KK); break;                                                       \
    case 3: ZF_DB_LAUNCH(3, KK); break;                                                       \
    default: ZF_DB_LAUNCH(4, KK); break;                                                      \
  }
  if (h->kind == zf::DK_SEP) { ZF_DB_LAUNCH_R(zf::DK_SEP) }
  else if (h->kind == zf::DK_SYM) { ZF_DB_LAUNCH_R(zf::DK_SYM) }
  else { ZF_DB_LAUNCH_R(zf::DK_GENERAL) }
#undef ZF_DB_LAUNCH_R
#undef ZF_DB_LAUNCH
  ZF_CUDA(cudaGetLastError());
  if (!h->capturing) zf::zf_count_launch();
  return ZF_OK;
}

int launch_prox_g(zf_deblur* h, const RunGroup& g, const zf::DeblurCtl& c) {
  dim3 grid((unsigned)h->d.prox_blocks, (unsigned)g.n_runs);
  zf::DeblurBufs B = h->B;
  B.run0 = g.run0;
  B.group = g.group;
  zf::deblur_prox_kernel<<<grid, zf::DB_THREADS, 0, g.st>>>(h->d, c, B);
  ZF_CUDA(cudaGetLastError());
  if (!h->capturing) zf::zf_count_launch();
  return ZF_OK;
}

template <int MODE>
int launch_tile(zf_deblur* h, int n_runs, const zf::DeblurCtl& c, bool decide, bool count) {
  return launch_tile_g<MODE>(h, RunGroup{h->st, 0, n_runs, 0}, c, decide, count);
}

int deblur_check_options(const zf_options* o) {
  if (!o) return zf::zf_fail(ZF_ERR_INVALID, "options is NULL");
  if (!(o->lr > 0.0)) return zf::zf_fail(ZF_ERR_INVALID, "lr must be > 0");
  if (o->max_iter < 1) return zf::zf_fail(ZF_ERR_INVALID, "max_iter must be >= 1");
  if (o->max_backtrack_iter < 1) return zf::zf_fail(ZF_ERR_INVALID, "max_backtrack_iter must be >= 1");
  if (!(o->decay_rate > 0.0 && o->decay_rate <= 1.0))
    return zf::zf_fail(ZF_ERR_INVALID, "decay_rate must be in (0, 1]");
  if (o->trace_capacity < 0) return zf::zf_fail(ZF_ERR_INVALID, "trace_capacity must be >= 0");
  return ZF_OK;
}

int zero_partials(zf_deblur* h, int n_runs) {
  const size_t part_bytes = sizeof(double) * (size_t)n_runs * h->d.n_tiles;
  ZF_CUDA(cudaMemsetAsync(h->B.fy_part, 0, part_bytes, h->st));
  ZF_CUDA(cudaMemsetAsync(h->B.fx_part, 0, part_bytes, h->st));
  ZF_CUDA(cudaMemsetAsync(h->B.abs_part, 0, part_bytes, h->st));
  ZF_CUDA(cudaMemsetAsync(h->B.tsum, 0, sizeof(zf::DeblurSums) * (size_t)n_runs * h->d.n_tiles, h->st));
  ZF_CUDA(cudaMemsetAsync(h->B.psum, 0, sizeof(zf::DeblurSums) * (size_t)n_runs * h->d.prox_blocks, h->st));
  ZF_CUDA(cudaMemsetAsync(h->B.tickets, 0, sizeof(unsigned int) * (size_t)n_runs, h->st));
  return ZF_OK;
}

// Solve n_runs runs whose x0 is already in all three X buffers on the device.
int deblur_run(zf_deblur* h, const zf_options* opt, int n_runs, const double* h_ab,
               bool want_funs) {
  const int cap = opt->trace_capacity;
  zf::DeblurCtl c{};
  c.tol = opt->tol; c.tol_internal = opt->tol_internal; c.decay_rate = opt->decay_rate;
  c.l1 = h->l1; c.max_iter = opt->max_iter; c.max_backtrack = opt->max_backtrack_iter;
  c.nesterov = opt->nesterov; c.deprecated = opt->deprecated; c.cap = cap;
  c.need_F = (opt->decay_rate != 1.0) || (cap > 0 && want_funs);
  c.store_yg = (opt->decay_rate != 1.0);
  if (cap > 0) {
    const size_t need = (size_t)n_runs * ((size_t)cap + 1);
    if (need > h->trace_cap_alloc) {
      cudaFree(h->B.allerrs); cudaFree(h->B.allfuns);
      h->B.allerrs = h->B.allfuns = nullptr;
      h->trace_cap_alloc = 0;
      ZF_CUDA(cudaMalloc((void**)&h->B.allerrs, need * 8));
      ZF_CUDA(cudaMalloc((void**)&h->B.allfuns, need * 8));
      h->trace_cap_alloc = need;
    }
    ZF_CUDA(cudaMemsetAsync(h->B.allerrs, 0, (size_t)n_runs * cap * 8, h->st));
    ZF_CUDA(cudaMemsetAsync(h->B.allfuns, 0, (size_t)n_runs * (cap + 1) * 8, h->st));
  }
  std::vector<zf::DeblurRun> init((size_t)n_runs);
  for (int r = 0; r < n_runs; ++r) {
    zf::DeblurRun& s = init[r];
    std::memset(&s, 0, sizeof(s));
    s.lr = opt->lr; s.t_prev = 1.0; s.mom = 0.0;
    s.na = h_ab ? h_ab[2 * r] : opt->nesterov_a;
    s.nb = h_ab ? h_ab[2 * r + 1] : opt->nesterov_b;
    s.err = INFINITY;
    s.nit = c.need_F ? 0 : 1;
    s.phase = c.need_F ? zf::DP_INIT : zf::DP_NEW;
    s.cur = 0; s.prev = 1; s.nxt = 2; s.res_buf = 0;
  }
  ZF_CUDA(cudaMemcpyAsync(h->B.runs, init.data(), sizeof(zf::DeblurRun) * n_runs,
                          cudaMemcpyHostToDevice, h->st));
  int rc = zero_partials(h, n_runs);
  if (rc != ZF_OK) return rc;
  ZF_CUDA(cudaStreamSynchronize(h->st));   // `init` must outlive the copy
  // The runs are split into two groups that advance on two streams.  Runs are independent,
  // and a round is one wave of CTAs whose phases (gather from L2, then two FP64 stencils) all
  // CTAs go through in lockstep; two groups drift apart, so one group's gather overlaps the
  // other's stencils.
  int n_groups = (n_runs >= 4) ? 2 : 1;
  if (const char* env = getenv("ZF_DEBLUR_GROUPS")) {
    const int v = atoi(env);
    if (v >= 1 && v <= 4) n_groups = v < n_runs ? v : n_runs;
  }
  RunGroup groups[4];
  for (int gi = 0, r0 = 0; gi < n_groups; ++gi) {
    const int cnt = (n_runs - r0 + (n_groups - gi) - 1) / (n_groups - gi);
    groups[gi] = RunGroup{gi == 0 ? h->st : h->st2[gi - 1], r0, cnt, gi};
    r0 += cnt;
  }
  if (n_groups > 1) {
    ZF_CUDA(cudaEventRecord(h->ev, h->st));
    for (int gi = 1; gi < n_groups; ++gi) ZF_CUDA(cudaStreamWaitEvent(groups[gi].st, h->ev, 0));
  }
  // one round of a group; `count`: the deciding kernel also counts the runs that stay active
  auto round = [&](const RunGroup& g, bool count) -> int {
    int r2;
    if (count) ZF_CUDA(cudaMemsetAsync(h->B.n_active + g.group, 0, sizeof(unsigned int), g.st));
    if (!c.need_F) return launch_tile_g<0>(h, g, c, true, count);
    if ((r2 = launch_tile_g<0>(h, g, c, false, false)) != ZF_OK) return r2;
    if (c.store_yg && (r2 = launch_prox_g(h, g, c)) != ZF_OK) return r2;
    return launch_tile_g<1>(h, g, c, true, count);
  };
  // Chunks of rounds between polls (8, 16, 32, 64, 64, ...: short solves do not over-run much,
  // long ones poll rarely).  A chunk is captured once into a CUDA graph and replayed: every
  // kernel argument is constant during a solve (all state is in device memory).
  const int kernels_per_round = c.need_F ? (c.store_yg ? 3 : 2) : 1;
  cudaGraphExec_t execs[4][4] = {};
  auto build_graph = [&](const RunGroup& g, int chunk, cudaGraphExec_t* out) -> int {
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(g.st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();
      return ZF_ERR_CUDA;
    }
    h->capturing = true;
    int r2 = ZF_OK;
    for (int k = 0; k < chunk && r2 == ZF_OK; ++k) r2 = round(g, k == chunk - 1);
    if (r2 == ZF_OK &&
        cudaMemcpyAsync(h->h_active + g.group, h->B.n_active + g.group, sizeof(unsigned int),
                        cudaMemcpyDeviceToHost, g.st) != cudaSuccess)
      r2 = ZF_ERR_CUDA;
    h->capturing = false;
    cudaError_t e = cudaStreamEndCapture(g.st, &graph);
    if (r2 != ZF_OK || e != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      return ZF_ERR_CUDA;
    }
    e = cudaGraphInstantiate(out, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { cudaGetLastError(); *out = nullptr; return ZF_ERR_CUDA; }
    return ZF_OK;
  };
  int chunk = 8, slot = 0;
  bool use_graph = true;
  bool active[4] = {n_groups > 0, n_groups > 1, n_groups > 2, n_groups > 3};
  rc = ZF_OK;
  while (rc == ZF_OK && (active[0] || active[1] || active[2] || active[3])) {
    for (int gi = 0; gi < n_groups && rc == ZF_OK; ++gi) {
      if (!active[gi]) continue;
      const RunGroup& g = groups[gi];
      if (use_graph && !execs[gi][slot] && build_graph(g, chunk, &execs[gi][slot]) != ZF_OK)
        use_graph = false;
      if (use_graph) {
        if (cudaGraphLaunch(execs[gi][slot], g.st) != cudaSuccess) rc = zf::zf_fail(ZF_ERR_CUDA, "cudaGraphLaunch failed");
        else zf::zf_count_launch((int64_t)kernels_per_round * chunk);
      } else {
        for (int k = 0; k < chunk && rc == ZF_OK; ++k) rc = round(g, k == chunk - 1);
        if (rc == ZF_OK &&
            cudaMemcpyAsync(h->h_active + gi, h->B.n_active + gi, sizeof(unsigned int),
                            cudaMemcpyDeviceToHost, g.st) != cudaSuccess)
          rc = zf::zf_fail(ZF_ERR_CUDA, "memcpy failed");
      }
    }
    for (int gi = 0; gi < n_groups && rc == ZF_OK; ++gi) {
      if (!active[gi]) continue;
      if (cudaStreamSynchronize(groups[gi].st) != cudaSuccess) rc = zf::zf_fail(ZF_ERR_CUDA, "stream sync failed");
      else if (h->h_active[gi] == 0) active[gi] = false;
    }
    if (chunk < 64) { chunk *= 2; ++slot; }
  }
  for (auto& row : execs)
    for (auto& ex : row)
      if (ex) cudaGraphExecDestroy(ex);
  if (rc != ZF_OK) return rc;
  if (!c.need_F) {   // res.fun = F(x): one evaluation of the result buffer
    c.finalize = 1;
    if ((rc = launch_tile<1>(h, n_runs, c, true, false)) != ZF_OK) return rc;
  }
  return ZF_OK;
}
}  // namespace

extern "C" int zf_deblur_create(zf_deblur** out, int32_t height, int32_t width,
                                const double* h_kernel, int32_t ksize, const double* h_observed,
                                double l1, int32_t max_runs, void* cuda_stream) {
  if (!out || !h_kernel || !h_observed) return zf::zf_fail(ZF_ERR_INVALID, "NULL argument");
  if (ksize < 3 || ksize > 9 || ksize % 2 == 0)
    return zf::zf_fail(ZF_ERR_UNSUPPORTED, "kernel size must be odd, 3..9 (got %d)", ksize);
  if (height < 2 * ksize || width < 2 * ksize || height % 2 || width % 2)
    return zf::zf_fail(ZF_ERR_INVALID, "image sides must be even and >= 2*ksize");
  if (max_runs < 1) return zf::zf_fail(ZF_ERR_INVALID, "max_runs must be >= 1");
  int rc = zf::zf_require_device();
  if (rc != ZF_OK) return rc;
  zf_deblur* h = new zf_deblur();
  zf::DeblurDims& d = h->d;
  d.H = height; d.W = width; d.h2 = height / 2; d.w2 = width / 2;
  d.tiles_x = (width + zf::DB_T - 1) / zf::DB_T;
  d.tiles_y = (height + zf::DB_T - 1) / zf::DB_T;
  d.n_tiles = d.tiles_x * d.tiles_y;
  d.R = ksize / 2;
  d.n = (long long)height * width;
  long long pb = (d.n + zf::DB_THREADS * 4 - 1) / (zf::DB_THREADS * 4);
  d.prox_blocks = (int)(pb > 256 ? 256 : pb);
  h->l1 = l1;
  h->max_runs = max_runs;
  h->st = (cudaStream_t)cuda_stream;
  if (!h->st) {
    if (cudaStreamCreate(&h->own_st) != cudaSuccess) {
      delete h;
      return zf::zf_fail(ZF_ERR_CUDA, "cudaStreamCreate failed");
    }
    h->st = h->own_st;
  }
  cudaGetDevice(&h->device);
  std::memcpy(h->taps.k, h_kernel, sizeof(double) * ksize * ksize);
  {
    // which stencil form.  Column symmetry is tested exactly.  Separability: with the centre row
    // and column as factors, a[u] = k[u][c] / sqrt(k[c][c]), b[v] = k[c][v] / sqrt(k[c][c]), every
    // tap must be a[u] b[v] to within a few ulps of the largest tap (an outer product that was
    // normalised afterwards, as in examples/cameraman.ipynb, is rank one only up to rounding; the
    // separable operator then differs from the given one by ~1e-16 relative, far below the 1e-8
    // parity tolerance).  ZF_DEBLUR_FORM=general|sym|sep overrides (sym / sep only if they apply).
    bool sym = true;
    for (int u = 0; u < ksize && sym; ++u)
      for (int v = 0; v < ksize / 2; ++v)
        if (h_kernel[u * ksize + v] != h_kernel[u * ksize + (ksize - 1 - v)]) { sym = false; break; }
    const int cc = ksize / 2;
    const double centre = h_kernel[cc * ksize + cc];
    bool sep = centre > 0.0;
    if (sep) {
      const double root = std::sqrt(centre);
      double kmax = 0.0;
      for (int i = 0; i < ksize * ksize; ++i) kmax = std::fmax(kmax, std::fabs(h_kernel[i]));
      for (int u = 0; u < ksize; ++u) {
        h->taps.a[u] = h_kernel[u * ksize + cc] / root;
        h->taps.b[u] = h_kernel[cc * ksize + u] / root;
      }
      for (int u = 0; u < ksize && sep; ++u)
        for (int v = 0; v < ksize; ++v)
          if (std::fabs(h->taps.a[u] * h->taps.b[v] - h_kernel[u * ksize + v]) > 8 * 2.3e-16 * kmax) {
            sep = false;
            break;
          }
    }
    h->kind = sep ? zf::DK_SEP : (sym ? zf::DK_SYM : zf::DK_GENERAL);
    if (const char* env = getenv("ZF_DEBLUR_FORM")) {
      if (env[0] == 'g') h->kind = zf::DK_GENERAL;
      else if (env[0] == 's' && env[1] == 'y' && sym) h->kind = zf::DK_SYM;
      else if (env[0] == 's' && env[1] == 'e' && sep) h->kind = zf::DK_SEP;
    }
    if (const char* env = getenv("ZF_DEBLUR_SYM"))       // round-1 switch: 0 = never fold
      if (env[0] == '0' && h->kind == zf::DK_SYM) h->kind = zf::DK_GENERAL;
  }
  const size_t vb = sizeof(double) * (size_t)max_runs * (size_t)d.n;
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
  zf::DeblurBufs& B = h->B;
  double* bimg = nullptr;
  alloc((void**)&bimg, sizeof(double) * (size_t)d.n);
  B.b = bimg;
  for (int k = 0; k < 3; ++k) alloc((void**)&B.X[k], vb);
  alloc((void**)&B.Y, vb); alloc((void**)&B.G, vb); alloc((void**)&h->scratch, vb);
  alloc((void**)&B.fy_part, sizeof(double) * (size_t)max_runs * d.n_tiles);
  alloc((void**)&B.fx_part, sizeof(double) * (size_t)max_runs * d.n_tiles);
  alloc((void**)&B.abs_part, sizeof(double) * (size_t)max_runs * d.n_tiles);
  alloc((void**)&B.tsum, sizeof(zf::DeblurSums) * (size_t)max_runs * d.n_tiles);
  alloc((void**)&B.psum, sizeof(zf::DeblurSums) * (size_t)max_runs * d.prox_blocks);
  alloc((void**)&B.runs, sizeof(zf::DeblurRun) * (size_t)max_runs);
  alloc((void**)&B.tickets, sizeof(unsigned int) * (size_t)max_runs);
  alloc((void**)&B.n_active, 4 * sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMallocHost((void**)&h->h_active, 4 * sizeof(unsigned int));
  for (int k = 0; k < 3; ++k)
    if (e == cudaSuccess) e = cudaStreamCreate(&h->st2[k]);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev, cudaEventDisableTiming);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(bimg, h_observed, sizeof(double) * (size_t)d.n, cudaMemcpyHostToDevice, h->st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
  if (e != cudaSuccess) {
    zf_deblur_destroy(h);
    return zf::zf_fail_cuda(e, "zf_deblur_create");
  }
  *out = h;
  return ZF_OK;
}

extern "C" void zf_deblur_destroy(zf_deblur* h) {
  if (!h) return;
  zf::DeblurBufs& B = h->B;
  cudaFree(const_cast<double*>(B.b));
  for (int k = 0; k < 3; ++k) cudaFree(B.X[k]);
  cudaFree(B.Y); cudaFree(B.G); cudaFree(h->scratch);
  cudaFree(B.fy_part); cudaFree(B.fx_part); cudaFree(B.abs_part); cudaFree(B.tsum);
  cudaFree(B.psum); cudaFree(B.runs); cudaFree(B.tickets); cudaFree(B.n_active);
  cudaFree(B.allerrs); cudaFree(B.allfuns);
  if (h->h_active) cudaFreeHost(h->h_active);
  if (h->own_st) cudaStreamDestroy(h->own_st);
  for (int k = 0; k < 3; ++k)
    if (h->st2[k]) cudaStreamDestroy(h->st2[k]);
  if (h->ev) cudaEventDestroy(h->ev);
  delete h;
}

// a handle works on the device it was created on, whatever the caller's current device is
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev); else prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// x0: n_runs x n when x0_is_batched, else one vector shared by every run.
static int deblur_solve_impl(zf_deblur* h, const zf_options* opt, int64_t n_runs,
                             const double* x0, bool x0_on_device, int x0_is_batched,
                             const double* h_ab, const zf_result* out, bool out_on_device) {
  if (!h || !x0 || !out) return zf::zf_fail(ZF_ERR_INVALID, "NULL argument");
  int rc = deblur_check_options(opt);
  if (rc != ZF_OK) return rc;
  if (n_runs < 0 || n_runs > h->max_runs)
    return zf::zf_fail(ZF_ERR_INVALID, "n_runs=%lld outside 0..max_runs=%d", (long long)n_runs, h->max_runs);
  if (n_runs == 0) return ZF_OK;
  if (!out->x || !out->fun || !out->nit || !out->status)
    return zf::zf_fail(ZF_ERR_INVALID, "result.x/fun/nit/status are required");
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceGuard on_device(h->device);
  const size_t nb = sizeof(double) * (size_t)h->d.n;
  const cudaMemcpyKind kin = x0_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  for (int64_t r = 0; r < n_runs; ++r) {
    const double* src = x0 + (x0_is_batched ? (size_t)r * h->d.n : 0);
    ZF_CUDA(cudaMemcpyAsync(h->B.X[0] + (size_t)r * h->d.n, src, nb, kin, h->st));
  }
  // x^0 = x^{-1} = x0 (proximal_gradient.py:463-465); the third buffer is the first candidate
  ZF_CUDA(cudaMemcpyAsync(h->B.X[1], h->B.X[0], nb * (size_t)n_runs, cudaMemcpyDeviceToDevice, h->st));
  rc = deblur_run(h, opt, (int)n_runs, h_ab, out->allfuns != nullptr);
  if (rc != ZF_OK) return rc;
  // results
  const cudaMemcpyKind kout = out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  dim3 grid((unsigned)h->d.prox_blocks, (unsigned)n_runs);
  zf::deblur_gather_kernel<<<grid, zf::DB_THREADS, 0, h->st>>>(h->d, h->B, h->scratch);
  ZF_CUDA(cudaGetLastError());
  zf::zf_count_launch();
  ZF_CUDA(cudaMemcpyAsync(out->x, h->scratch, nb * (size_t)n_runs, kout, h->st));
  std::vector<zf::DeblurRun> fin((size_t)n_runs);
  ZF_CUDA(cudaMemcpyAsync(fin.data(), h->B.runs, sizeof(zf::DeblurRun) * n_runs, cudaMemcpyDeviceToHost, h->st));
  const int cap = opt->trace_capacity;
  if (cap > 0 && out->allerrs)
    ZF_CUDA(cudaMemcpyAsync(out->allerrs, h->B.allerrs, (size_t)n_runs * cap * 8, kout, h->st));
  if (cap > 0 && out->allfuns)
    ZF_CUDA(cudaMemcpyAsync(out->allfuns, h->B.allfuns, (size_t)n_runs * (cap + 1) * 8, kout, h->st));
  ZF_CUDA(cudaStreamSynchronize(h->st));
  if (out_on_device) {
    // scalars are small: stage on host, copy up
    std::vector<double> fun(n_runs), lr(n_runs), err(n_runs);
    std::vector<int64_t> nit(n_runs);
    std::vector<int32_t> status(n_runs);
    for (int64_t r = 0; r < n_runs; ++r) {
      fun[r] = fin[r].F_x; lr[r] = fin[r].lr; err[r] = fin[r].err; nit[r] = fin[r].nit; status[r] = fin[r].status;
    }
    ZF_CUDA(cudaMemcpyAsync(out->fun, fun.data(), 8 * n_runs, cudaMemcpyHostToDevice, h->st));
    ZF_CUDA(cudaMemcpyAsync(out->nit, nit.data(), 8 * n_runs, cudaMemcpyHostToDevice, h->st));
    ZF_CUDA(cudaMemcpyAsync(out->status, status.data(), 4 * n_runs, cudaMemcpyHostToDevice, h->st));
    if (out->lr) ZF_CUDA(cudaMemcpyAsync(out->lr, lr.data(), 8 * n_runs, cudaMemcpyHostToDevice, h->st));
    if (out->err) ZF_CUDA(cudaMemcpyAsync(out->err, err.data(), 8 * n_runs, cudaMemcpyHostToDevice, h->st));
    ZF_CUDA(cudaStreamSynchronize(h->st));
  } else {
    for (int64_t r = 0; r < n_runs; ++r) {
      out->fun[r] = fin[r].F_x;
      out->nit[r] = fin[r].nit;
      out->status[r] = fin[r].status;
      if (out->lr) out->lr[r] = fin[r].lr;
      if (out->err) out->err[r] = fin[r].err;
    }
  }
  return ZF_OK;
}

extern "C" int zf_deblur_solve_host(zf_deblur* h, const zf_options* opt, int64_t n_runs,
                                    const double* h_x0, int32_t x0_is_batched,
                                    const double* h_ab, const zf_result* h_out) {
  return deblur_solve_impl(h, opt, n_runs, h_x0, false, x0_is_batched, h_ab, h_out, false);
}

extern "C" int zf_deblur_solve_device(zf_deblur* h, const zf_options* opt, int64_t n_runs,
                                      const double* d_x0, int32_t x0_is_batched,
                                      const double* h_ab, const zf_result* d_out) {
  return deblur_solve_impl(h, opt, n_runs, d_x0, true, x0_is_batched, h_ab, d_out, true);
}

extern "C" int zf_deblur_eval_host(zf_deblur* h, int64_t n_points, const double* h_X,
                                   double* h_f, double* h_g, double* h_jac) {
  if (!h || !h_X) return zf::zf_fail(ZF_ERR_INVALID, "NULL argument");
  if (n_points < 0 || n_points > h->max_runs)
    return zf::zf_fail(ZF_ERR_INVALID, "n_points outside 0..max_runs");
  if (n_points == 0) return ZF_OK;
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceGuard on_device(h->device);
  int rc = ZF_OK;
  const int n = (int)n_points;
  const size_t nb = sizeof(double) * (size_t)h->d.n * n;
  ZF_CUDA(cudaMemcpyAsync(h->B.X[0], h_X, nb, cudaMemcpyHostToDevice, h->st));
  ZF_CUDA(cudaMemcpyAsync(h->B.X[1], h_X, nb, cudaMemcpyHostToDevice, h->st));
  rc = zero_partials(h, n);
  if (rc != ZF_OK) return rc;
  zf::DeblurCtl c{};
  c.l1 = h->l1; c.decay_rate = 0.5; c.need_F = 1; c.store_yg = 1;
  std::vector<zf::DeblurRun> init((size_t)n);
  for (auto& s : init) {
    std::memset(&s, 0, sizeof(s));
    s.phase = zf::DP_NEW; s.lr = 1.0; s.cur = 0; s.prev = 1; s.nxt = 2;
  }
  ZF_CUDA(cudaMemcpyAsync(h->B.runs, init.data(), sizeof(zf::DeblurRun) * n, cudaMemcpyHostToDevice, h->st));
  rc = launch_tile<0>(h, n, c, false, false);      // G = jac_f(x) (no state change)
  if (rc != ZF_OK) return rc;
  for (auto& s : init) s.phase = zf::DP_INIT;
  ZF_CUDA(cudaStreamSynchronize(h->st));
  ZF_CUDA(cudaMemcpyAsync(h->B.runs, init.data(), sizeof(zf::DeblurRun) * n, cudaMemcpyHostToDevice, h->st));
  rc = launch_tile<1>(h, n, c, false, false);      // fx_part, abs_part of x
  if (rc != ZF_OK) return rc;
  std::vector<double> fp((size_t)n * h->d.n_tiles), ap((size_t)n * h->d.n_tiles);
  ZF_CUDA(cudaMemcpyAsync(fp.data(), h->B.fx_part, fp.size() * 8, cudaMemcpyDeviceToHost, h->st));
  ZF_CUDA(cudaMemcpyAsync(ap.data(), h->B.abs_part, ap.size() * 8, cudaMemcpyDeviceToHost, h->st));
  if (h_jac) ZF_CUDA(cudaMemcpyAsync(h_jac, h->B.G, nb, cudaMemcpyDeviceToHost, h->st));
  ZF_CUDA(cudaStreamSynchronize(h->st));
  for (int r = 0; r < n; ++r) {
    // same fixed order as deblur_decide: lane-strided partial sums, then butterfly
    double lf[32] = {0}, la[32] = {0};
    for (int t = 0; t < h->d.n_tiles; ++t) { lf[t & 31] += fp[(size_t)r * h->d.n_tiles + t]; la[t & 31] += ap[(size_t)r * h->d.n_tiles + t]; }
    for (int o = 16; o > 0; o >>= 1)
      for (int l = 0; l < o; ++l) { lf[l] += lf[l + o]; la[l] += la[l + o]; }
    const double nrm = std::sqrt(lf[0]);
    if (h_f) h_f[r] = nrm * nrm;
    if (h_g) h_g[r] = h->l1 * la[0];
  }
  return ZF_OK;
}
