// Large-n single-objective LASSO path:  F(x) = scale*||A x - b||^2 + l1*||x||_1
// with dense row-major fp64 A streamed from HBM (north_star (c); replaces
// minimize_proximal_gradient(f, g, jac_f, prox_wsum_g, x0, ...) driven by the dense numpy
// closures of tests/test_proximal_gradient.py:49-63 at sizes where A >> L2).
//
// Per FISTA iteration the device does
//   q = A^T (A y - b), sum r^2       ONE pass over A with a fused kernel chosen per shape
//                                    (lasso_fused_ring_kernel: warp-specialised TMA chunk
//                                     ring, the default from 4096 columns up;
//                                     lasso_fused_kernel: single CTA, L2 re-read, below that)
//                                    or two passes where a row is too wide for them:
//                                    lasso_residual_kernel (r = A y - b), lasso_atr_kernel (A^T r)
//   x = soft(y - lr*2*scale*q, lr*l1) + the four sums the line search / stop test need,
//   collect of the row-block partials, y = x + mom*(x - x_prev), stop test, t_{k+1}
//                                    lasso_dev_update_kernel (n_cols work, ONE kernel)
//   [line search]  ||A x - b||^2     lasso_residual_kernel   (one pass per trial), then
//                                    lasso_dev_decide_kernel + lasso_dev_momentum_kernel
// Every reduction has a fixed order (no floating-point atomics), so a solve is bit
// reproducible run to run.  Scalars (lr, t_k, F values, accept / stop decisions) live in DEVICE
// memory (LassoDevState): zf_lasso_solve enqueues trials as CUDA graphs of 32 and polls a flag
// one chunk behind; the round-1 host-decided loop (zf_lasso_begin / grad / step / finish, one
// D2H copy + stream sync per trial) is kept as the split API and the A/B baseline.
//
// Row-sharded multi-GPU: each rank owns a block of rows and the same replicated vectors; the
// only exchange is the sum over ranks of `partial` = [A^T r (n_cols) | sum r^2], done INSIDE the
// update kernel by reading every rank's published partials from NVLink peer memory (P2PBuf,
// zf_lasso_p2p_*), or by the caller's NCCL all-reduce between the stages.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cooperative_groups.h>

#include "zf_common.cuh"
#include "zf_host.h"

namespace zf {

// streaming 16 B load that does not displace the (reused) vectors from L1
__device__ __forceinline__ double2 ld_stream2(const double* p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];"
               : "=d"(v.x), "=d"(v.y)
               : "l"(p));
  return v;
}
__device__ __forceinline__ double ld_stream1(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

constexpr int RES_THREADS = 256;
constexpr int RES_ROWS_PER_WARP = 4;

// r = A v - b for a block of rows per warp; sq_part[block] = sum over the block's rows of r^2.
// A warp owns RES_ROWS_PER_WARP consecutive rows at a time: each v element it loads is used
// for all of them, and the independent row streams give 16 outstanding 16 B loads per lane.
template <bool VEC>
__global__ void __launch_bounds__(RES_THREADS, 3)
lasso_residual_kernel(const double* __restrict__ A, const double* __restrict__ b,
                      const double* __restrict__ v, long long n_rows, long long n_cols,
                      double* __restrict__ r, double* __restrict__ sq_part,
                      const int* __restrict__ skip) {
  if (skip != nullptr && *reinterpret_cast<const volatile int*>(skip) != 0) return;   // device-decided loop: nothing to do
  constexpr int R = RES_ROWS_PER_WARP;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  constexpr int WARPS = RES_THREADS / 32;
  const long long n_groups = (n_rows + R - 1) / R;
  const long long total_warps = (long long)gridDim.x * WARPS;
  double ss = 0.0;
  for (long long g = (long long)blockIdx.x * WARPS + warp; g < n_groups; g += total_warps) {
    const long long row0 = g * R;
    const double* rowp[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const long long row = (row0 + k < n_rows) ? row0 + k : n_rows - 1;   // clamp, masked below
      rowp[k] = A + row * n_cols;
    }
    double acc[R];
#pragma unroll
    for (int k = 0; k < R; ++k) acc[k] = 0.0;
    if (VEC) {
      const long long n2 = n_cols >> 1;
      long long c = lane;
      for (; c + 96 < n2; c += 128) {       // 4 column steps x R rows = 16 loads in flight
        double2 a[4][R];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int k = 0; k < R; ++k) a[u][k] = ld_stream2(rowp[k] + 2 * (c + 32 * u));
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const double2 vv = __ldg(reinterpret_cast<const double2*>(v) + c + 32 * u);
#pragma unroll
          for (int k = 0; k < R; ++k) acc[k] += a[u][k].x * vv.x + a[u][k].y * vv.y;
        }
      }
      for (; c < n2; c += 32) {
        const double2 vv = __ldg(reinterpret_cast<const double2*>(v) + c);
#pragma unroll
        for (int k = 0; k < R; ++k) {
          const double2 a = ld_stream2(rowp[k] + 2 * c);
          acc[k] += a.x * vv.x + a.y * vv.y;
        }
      }
    } else {
      for (long long c = lane; c < n_cols; c += 32) {
        const double vv = __ldg(v + c);
#pragma unroll
        for (int k = 0; k < R; ++k) acc[k] += ld_stream1(rowp[k] + c) * vv;
      }
    }
    warp_sum_k<R>(acc);
#pragma unroll
    for (int k = 0; k < R; ++k) {
      if (row0 + k < n_rows) {
        const double d = acc[k] - b[row0 + k];
        if (lane == 0) r[row0 + k] = d;
        ss += d * d;
      }
    }
  }
  __shared__ double wsum[WARPS];
  if (lane == 0) wsum[warp] = ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) t += wsum[w];
    sq_part[blockIdx.x] = t;
  }
}

constexpr int ATR_THREADS = 256;
constexpr int ATR_COLS_PER_THREAD = 4;                       // two 16 B column pairs
constexpr int ATR_SLAB = ATR_THREADS * ATR_COLS_PER_THREAD;  // 1024 columns per CTA
constexpr int ATR_ROW_UNROLL = 8;

// gpart[rb][j] = sum over the rows of row-block rb of r_i * A[i][j].  CTA (slab, rb): threads own
// columns (accumulators in registers), rows stream through; 16 outstanding 16 B loads/thread.
template <bool VEC>
__global__ void __launch_bounds__(ATR_THREADS, 2)
lasso_atr_kernel(const double* __restrict__ A, const double* __restrict__ r, long long n_rows,
                 long long n_cols, long long rows_per_block, double* __restrict__ gpart,
                 const int* __restrict__ skip) {
  if (skip != nullptr && *reinterpret_cast<const volatile int*>(skip) != 0) return;   // device-decided loop: nothing to do
  const long long col_base = (long long)blockIdx.x * ATR_SLAB;
  const long long rb = blockIdx.y;
  const long long i0 = rb * rows_per_block;
  const long long i1 = (i0 + rows_per_block < n_rows) ? i0 + rows_per_block : n_rows;
  double* out = gpart + rb * n_cols;
  if (VEC) {
    const long long c0 = col_base + 2 * threadIdx.x;             // first pair
    const long long c1 = c0 + 2 * ATR_THREADS;                   // second pair
    const bool ok0 = c0 < n_cols, ok1 = c1 < n_cols;
    // clamp addresses of inactive pairs onto a valid one; their results are never stored
    const long long a0 = ok0 ? c0 : 0, a1 = ok1 ? c1 : 0;
    double2 s0 = make_double2(0.0, 0.0), s1 = make_double2(0.0, 0.0);
    long long i = i0;
    for (; i + ATR_ROW_UNROLL <= i1; i += ATR_ROW_UNROLL) {
      double2 p[ATR_ROW_UNROLL], q[ATR_ROW_UNROLL];
#pragma unroll
      for (int u = 0; u < ATR_ROW_UNROLL; ++u) {
        const double* row = A + (i + u) * n_cols;
        p[u] = ld_stream2(row + a0);
        q[u] = ld_stream2(row + a1);
      }
#pragma unroll
      for (int u = 0; u < ATR_ROW_UNROLL; ++u) {
        const double ri = __ldg(r + i + u);
        s0.x += ri * p[u].x; s0.y += ri * p[u].y;
        s1.x += ri * q[u].x; s1.y += ri * q[u].y;
      }
    }
    for (; i < i1; ++i) {
      const double* row = A + i * n_cols;
      const double2 p = ld_stream2(row + a0), q = ld_stream2(row + a1);
      const double ri = __ldg(r + i);
      s0.x += ri * p.x; s0.y += ri * p.y;
      s1.x += ri * q.x; s1.y += ri * q.y;
    }
    if (ok0) *reinterpret_cast<double2*>(out + c0) = s0;
    if (ok1) *reinterpret_cast<double2*>(out + c1) = s1;
  } else {
    // scalar fallback (odd n_cols or unaligned A): thread owns columns tid + k*256
    double s[ATR_COLS_PER_THREAD];
    long long cj[ATR_COLS_PER_THREAD];
    bool ok[ATR_COLS_PER_THREAD];
#pragma unroll
    for (int k = 0; k < ATR_COLS_PER_THREAD; ++k) {
      cj[k] = col_base + threadIdx.x + (long long)k * ATR_THREADS;
      ok[k] = cj[k] < n_cols;
      if (!ok[k]) cj[k] = 0;
      s[k] = 0.0;
    }
    for (long long i = i0; i < i1; ++i) {
      const double* row = A + i * n_cols;
      const double ri = __ldg(r + i);
#pragma unroll
      for (int k = 0; k < ATR_COLS_PER_THREAD; ++k) s[k] += ri * ld_stream1(row + cj[k]);
    }
#pragma unroll
    for (int k = 0; k < ATR_COLS_PER_THREAD; ++k)
      if (ok[k]) out[cj[k]] = s[k];
  }
}

// ------------------------------------------------------------------------------------
// Fused gradient pass: A^T (A v - b) and sum r^2 with A read from HBM ONCE.
// One persistent CTA per SM owns a contiguous block of rows and every column: v sits in
// shared memory (8 * n_cols bytes), the CTA's partial A^T r sits in registers (each thread
// owns PAIRS column pairs), and rows stream through two at a time:
//   phase 1: load the two rows (HBM -> L2 -> registers), dot them with v, block-reduce;
//   phase 2: load the same two rows again -- now L2 hits, marked evict-first -- and
//            accumulate r_i * A[i][:] into the thread's columns.
// HBM traffic is one pass over A; the second read is served by the 126 MB L2 (148 CTAs x 2
// rows x 8 n_cols bytes = 47 MB in flight for n_cols = 20000).
// ------------------------------------------------------------------------------------
constexpr int FUSED_THREADS = 512;
constexpr int FUSED_ROWS = 2;
constexpr int FUSED_CHUNK = 4;     // column pairs loaded per thread per step (x2 rows in flight)

// L2 residency control through cache-hint policies (the plain .L2::evict_* qualifiers are
// only accepted on 32-byte loads)
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ double2 ld_hint2(const double* p, unsigned long long pol) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;"
               : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
  return v;
}

template <int PAIRS>
__global__ void __launch_bounds__(FUSED_THREADS, 1)
lasso_fused_kernel(const double* __restrict__ A, const double* __restrict__ b,
                   const double* __restrict__ v, long long n_rows, long long n_cols,
                   long long rows_per_cta, double* __restrict__ gpart,
                   double* __restrict__ sq_part,
    const int* __restrict__ skip) {
  if (skip != nullptr && *reinterpret_cast<const volatile int*>(skip) != 0) return;   // device-decided loop: nothing to do
  extern __shared__ double vsm[];                       // n_cols doubles
  __shared__ double red[2][FUSED_THREADS / 32][FUSED_ROWS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long n2 = n_cols >> 1;
  for (long long p = tid; p < n2; p += FUSED_THREADS)
    reinterpret_cast<double2*>(vsm)[p] = __ldg(reinterpret_cast<const double2*>(v) + p);
  __syncthreads();
  const unsigned long long keep_pol = l2_policy_evict_last();     // phase 1: stay in L2
  const unsigned long long last_pol = l2_policy_evict_first();    // phase 2: last use
  double2 q[PAIRS];
#pragma unroll
  for (int k = 0; k < PAIRS; ++k) q[k] = make_double2(0.0, 0.0);
  const long long i0 = (long long)blockIdx.x * rows_per_cta;
  const long long i1 = (i0 + rows_per_cta < n_rows) ? i0 + rows_per_cta : n_rows;
  double ss = 0.0;
  int parity = 0;
  for (long long i = i0; i < i1; i += FUSED_ROWS) {
    const bool two = (i + 1 < i1);
    const double* row0 = A + i * n_cols;
    const double* row1 = two ? row0 + n_cols : row0;     // a lone last row is read twice, used once
    // ---- phase 1: dot products
    double acc[FUSED_ROWS] = {0.0, 0.0};
#pragma unroll
    for (int k0 = 0; k0 < PAIRS; k0 += FUSED_CHUNK) {
      double2 a0[FUSED_CHUNK], a1[FUSED_CHUNK];
#pragma unroll
      for (int u = 0; u < FUSED_CHUNK; ++u) {
        const long long p = tid + (long long)(k0 + u) * FUSED_THREADS;
        const long long pc = (k0 + u < PAIRS && p < n2) ? p : 0;
        a0[u] = ld_hint2(row0 + 2 * pc, keep_pol);
        a1[u] = ld_hint2(row1 + 2 * pc, keep_pol);
      }
#pragma unroll
      for (int u = 0; u < FUSED_CHUNK; ++u) {
        const long long p = tid + (long long)(k0 + u) * FUSED_THREADS;
        if (k0 + u < PAIRS && p < n2) {
          const double2 vv = reinterpret_cast<const double2*>(vsm)[p];
          acc[0] += a0[u].x * vv.x + a0[u].y * vv.y;
          acc[1] += a1[u].x * vv.x + a1[u].y * vv.y;
        }
      }
    }
    warp_sum_k<FUSED_ROWS>(acc);
    if (lane == 0) { red[parity][warp][0] = acc[0]; red[parity][warp][1] = acc[1]; }
    __syncthreads();
    double r0 = 0.0, r1 = 0.0;
#pragma unroll
    for (int w = 0; w < FUSED_THREADS / 32; ++w) { r0 += red[parity][w][0]; r1 += red[parity][w][1]; }
    parity ^= 1;
    r0 -= b[i];
    r1 = two ? r1 - b[i + 1] : 0.0;
    ss += r0 * r0;
    ss += r1 * r1;
    // ---- phase 2: rank-1 updates of the thread's columns (rows re-read from L2)
#pragma unroll
    for (int k0 = 0; k0 < PAIRS; k0 += FUSED_CHUNK) {
      double2 a0[FUSED_CHUNK], a1[FUSED_CHUNK];
#pragma unroll
      for (int u = 0; u < FUSED_CHUNK; ++u) {
        const long long p = tid + (long long)(k0 + u) * FUSED_THREADS;
        const long long pc = (k0 + u < PAIRS && p < n2) ? p : 0;
        a0[u] = ld_hint2(row0 + 2 * pc, last_pol);
        a1[u] = ld_hint2(row1 + 2 * pc, last_pol);
      }
#pragma unroll
      for (int u = 0; u < FUSED_CHUNK; ++u) {
        if (k0 + u < PAIRS) {
          q[k0 + u].x += r0 * a0[u].x; q[k0 + u].y += r0 * a0[u].y;
          q[k0 + u].x += r1 * a1[u].x; q[k0 + u].y += r1 * a1[u].y;
        }
      }
    }
  }
  double* out = gpart + (long long)blockIdx.x * n_cols;
#pragma unroll
  for (int k = 0; k < PAIRS; ++k) {
    const long long p = tid + (long long)k * FUSED_THREADS;
    if (p < n2) reinterpret_cast<double2*>(out)[p] = q[k];
  }
  if (tid == 0) sq_part[blockIdx.x] = ss;
}

// ------------------------------------------------------------------------------------
// mbarrier / bulk-TMA helpers of the chunk-ring kernel below.  (Round 1 also had a 2-CTA cluster
// form that re-read the row pair from L2 and two TMA forms with one block reduction + one
// cluster barrier per stage, 0.65-0.77 of the one-pass bound; the ring kernel superseded them
// on every shape and they were removed in round 2 -- their numbers stay in DESIGN.md 3.3.)
// ------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes,
                                            unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// try_wait with a suspend-time hint: ptxas then emits TRYWAIT; @!p NANOSLEEP.SYNCS <ns>; PHASECHK --
// the warp sleeps until the barrier's phase completes (or the hint runs out) instead of spinning.
// Half of the ring kernel's 1.78 G warp instructions per pass at 100000 x 20000 were spin-loop
// instructions of waiting warps (ncu source page: 100 M TRYWAITs, ~8 instructions per trip);
// with the hint the kernel is 1 % faster at sustained clocks on every shape measured (0 = off).
#ifndef ZF_MBAR_HINT_NS
#define ZF_MBAR_HINT_NS 2000
#endif
// bounded wait: a protocol bug must trap, not hang the GPU (2^22 trips of at most 2 us)
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned done = 0;
  for (unsigned spin = 0; !done; ++spin) {
#if ZF_MBAR_HINT_NS > 0
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(smem_u32(bar)), "r"(parity), "r"((unsigned)ZF_MBAR_HINT_NS) : "memory");
#else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
#endif
    if (!done && spin > (ZF_MBAR_HINT_NS > 0 ? (1u << 22) : (1u << 26))) __trap();
  }
}

// ------------------------------------------------------------------------------------
// Fused one-pass gradient, warp-specialised chunk ring.
// The roles are split over warps and nothing waits for a barrier it does not need:
//   * producer warp: streams each row of the CTA's column slice as 16 KB
//     chunks (1-D bulk TMA) into a ring of RING_SLOTS chunks -- rows simply follow each other
//     through the ring, a slot is refilled as soon as the update warps release it, so HBM
//     requests never pause;
//   * dot warps (8): keep their part of v in REGISTERS, form the partial dot products of the
//     RPS rows of an exchange step chunk by chunk as the chunks land, reduce them in one
//     butterfly and hand the warp partials to
//   * the exchange warp: adds the 8 warp partials in warp order (rank 0 folds -b_i in) and
//     posts the CTA's part of every row of the step to every CTA of the cluster with st.async
//     (remote shared-memory store that completes on the remote `ready` mbarrier);
//   * update warps (8): wait for the step's `ready` mbarrier, sum the parts in rank order
//     and apply the rank-1 updates q += r * row from the SAME shared-memory chunks (q in
//     registers), releasing each chunk to the producer.
// The dot warps run up to a ring ahead of the update warps, so the cluster exchange latency
// is off the critical path.  Partial-dot slots and `ready` barriers are indexed by step modulo
// RING_NR >= 2 * (rows a ring can hold) + 2, which is what makes reuse race-free: a peer can
// post step n + RING_NR only after this CTA's update warps have consumed step n.
// Cluster sizes 1..8 (no power-of-two assumption); a row slice is at most 5 chunks: with 4
// chunks (8192 columns) the ring keeps three rows, with 5 (2.4 rows) the kernel drops to 0.64.
// RPS = 2 rows per step for slices of <= 3 chunks (a step's chunks stay in the ring until its
// update is done: two rows of 4-5 chunks would leave the producer no prefetch depth).
// ------------------------------------------------------------------------------------
constexpr int RING_GROUP = 256;                    // threads of the dot / update group
constexpr int RING_WARPS = RING_GROUP / 32;
constexpr int RING_THREADS = 2 * RING_GROUP + 64;  // 8 dot + 8 update warps, producer, exchange warp
// (five warps on a scheduler cap the kernel at 96 registers per thread.  Moving registers from
// the two service warps to the others with setmaxnreg -- 20 warps, 104 / 120 registers for the
// dot / update warps -- was measured and is not used: 0.93 instead of 0.98 at 16384 columns.)
constexpr int RING_CH_PAIRS = 1024;                // double2 per chunk: 2048 columns, 16 KB
constexpr int RING_SLOTS = 13;                    // 208 KB ring + 4-8 KB static: the most that fits in 227 KB
constexpr int RING_NR = 2 * RING_SLOTS + 2;
constexpr int RING_U = RING_CH_PAIRS / RING_GROUP; // double2 per thread and chunk
constexpr int RING_MAX_CLUSTER = 8;

__device__ __forceinline__ unsigned mapa_u32(unsigned addr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_local(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(unsigned long long* bar, unsigned parity) {
  unsigned done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return done != 0;
}
// 8-byte store into a (possibly remote) CTA of the cluster that completes on THAT CTA's
// mbarrier (async proxy, like a TMA write): the reader only needs a CTA-scope wait.  The first
// version used st.shared::cluster + mbarrier.arrive.release.cluster and a
// try_wait.acquire.cluster spin on the reader: every spin iteration compiled to CCTL.IVALL (an
// L1 invalidate) and the kernel ran at 0.43 of the roofline (22 % of the stall samples).
__device__ __forceinline__ void st_async_f64(unsigned remote_addr, double v, unsigned remote_bar) {
  asm volatile(
      "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
      ::"r"(remote_addr), "l"(__double_as_longlong(v)), "r"(remote_bar) : "memory");
}

// Timing experiment: compile with -DZF_RING_DEBUG=1 and run with ZF_LASSO_RING_DBG=1 to get
// clock64 stamps of rows 64..127 of CTA <ZF_LASSO_RING_DBG>: [0] first chunk issued, [1] dot warp 0 done, [6] all
// dot warps in, [2] part posted, [3] update warps saw `ready`, [4] last chunk released; dbg2 =
// per-dot-warp completion.  (The numbers quoted in DESIGN.md 3.3 come from these.)  The stamps
// are compiled out by default: their predicates were ~15 of the ~45 instructions an update
// warp spends per chunk.
#ifndef ZF_RING_DEBUG
#define ZF_RING_DEBUG 0
#endif
#ifndef ZF_RING_RPS_DEFAULT
#define ZF_RING_RPS_DEFAULT 2     // rows per exchange step for row slices of <= 3 chunks
#endif
__device__ long long zf_ring_dbg[8][64];
__device__ long long zf_ring_dbg2[8][64];
#if ZF_RING_DEBUG
#define RING_STAMP(k, row) \
  do { if (dbg && (int)blockIdx.x == dbg - 1 && (row) >= 64 && (row) < 128) zf_ring_dbg[k][(row) - 64] = clock64(); } while (0)
#define RING_STAMP2(w, row) \
  do { if (dbg && (int)blockIdx.x == dbg - 1 && (row) >= 64 && (row) < 128) zf_ring_dbg2[w][(row) - 64] = clock64(); } while (0)
#else
#define RING_STAMP(k, row) do { } while (0)
#define RING_STAMP2(w, row) do { } while (0)
#endif

template <int NCH, int RPS>
__global__ void __maxnreg__(96)   // 18 warps: 5 on one scheduler, whose register file holds 5 x 32 x 102
lasso_fused_ring_kernel(const double* __restrict__ A, const double* __restrict__ b,
                        const double* __restrict__ v, long long n_rows, long long n_cols,
                        long long rows_per_cluster, long long pairs_per_cta,
                        double* __restrict__ gpart, double* __restrict__ sq_part,
                        int dbg, long long row_begin, long long row_end, int part_base,
                        const int* __restrict__ skip) {
  if (skip != nullptr && *reinterpret_cast<const volatile int*>(skip) != 0) return;   // device-decided loop: nothing to do
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int csize = (int)cluster.num_blocks();
  const long long cid = blockIdx.x / csize;
  extern __shared__ __align__(128) unsigned char dyn[];
  // RPS rows per exchange STEP (2 for row slices of <= 3 chunks, else 1): the per-step fixed work
  // -- the warp reduction, two barrier operations, the ready wait, the sum of the parts -- is paid
  // once per step, and the dot warps run ahead by a whole row while the exchange is in flight
  // (sustained clocks, interleaved A/B: +2..14 % at <= 3 chunks per row)
  __shared__ double xslot[RING_NR][RPS][RING_MAX_CLUSTER];  // [step mod NR][row of the step][cluster rank]
  __shared__ double dpart[RING_NR][RPS][RING_WARPS];        // [step mod NR][row of the step][dot warp]
  __shared__ __align__(8) unsigned long long full[RING_SLOTS];
  __shared__ __align__(8) unsigned long long empty[RING_SLOTS];
  __shared__ __align__(8) unsigned long long ready[RING_NR];   // the row's residual parts are here
  __shared__ __align__(8) unsigned long long dbar[RING_NR];    // the row's 8 warp partials are here
  double2* ring = reinterpret_cast<double2*>(dyn);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long n2 = n_cols >> 1;
  const long long p_lo = (long long)crank * pairs_per_cta;
  const long long p_hi = (p_lo + pairs_per_cta < n2) ? p_lo + pairs_per_cta : n2;
  const int my_pairs = (int)(p_hi > p_lo ? p_hi - p_lo : 0);
  const int nch = (my_pairs + RING_CH_PAIRS - 1) / RING_CH_PAIRS;      // <= NCH
  if (tid == 0) {
    for (int s = 0; s < RING_SLOTS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], RING_WARPS);
    }
    for (int s = 0; s < RING_NR; ++s) {
      mbar_init(&ready[s], 1u);
      mbar_init(&dbar[s], RING_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster.sync();                                      // every CTA's barriers exist
  // this launch covers rows [row_begin, row_end) and writes partial rows part_base + cid
  const long long i0 = row_begin + cid * rows_per_cluster;
  const long long i1 = (i0 + rows_per_cluster < row_end) ? i0 + rows_per_cluster : row_end;
  const long long my_rows = i1 > i0 ? i1 - i0 : 0;

  if (warp == 2 * RING_WARPS) {
    // ---------------------------------------------------------------- producer
    // A dedicated warp: one 16 KB bulk copy occupies the issuing thread for ~600 cycles (measured
    // with clock64 stamps when a lane of an update warp issued them between its own work: the
    // update warps, and with them the whole ring, ran at the producer's pace -- 0.55 of the
    // roofline).  Five warps on one scheduler limit the kernel to 96 registers per thread.
    if (lane == 0) {
      int psl = 0, pc = 0;
      unsigned pph = 1;                                  // parity of the slot's previous release
      long long prow = 0;
      const long long total = my_rows * nch;
      for (long long jp = 0; jp < total; ++jp) {
        if (jp >= RING_SLOTS) mbar_wait(&empty[psl], pph);
        const int cp = (my_pairs - pc * RING_CH_PAIRS < RING_CH_PAIRS) ? my_pairs - pc * RING_CH_PAIRS
                                                                       : RING_CH_PAIRS;
        if (pc == 0) RING_STAMP(0, prow);
        mbar_expect_tx(&full[psl], (unsigned)cp * 16u);
        tma_load_1d(ring + (size_t)psl * RING_CH_PAIRS,
                    A + (i0 + prow) * n_cols + 2 * (p_lo + (long long)pc * RING_CH_PAIRS),
                    (unsigned)cp * 16u, &full[psl]);
        if (++pc == nch) { pc = 0; ++prow; }
        if (++psl == RING_SLOTS) { psl = 0; pph ^= 1u; }
      }
    }
  } else if (warp < RING_WARPS) {
    // ---------------------------------------------------------------- dot warps
    double2 vreg[NCH][RING_U];
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
      for (int u = 0; u < RING_U; ++u) {
        const int p = c * RING_CH_PAIRS + u * RING_GROUP + tid;
        vreg[c][u] = (p < my_pairs) ? __ldg(reinterpret_cast<const double2*>(v) + p_lo + p)
                                    : make_double2(0.0, 0.0);
      }
    int s = 0, sl = 0;
    unsigned ph = 0;                          // phase parity of the ring slot
    for (long long n = 0; n < my_rows; n += RPS) {
      double accu[RPS][RING_U];             // independent chains: a warp handles every chunk, its
#pragma unroll                              // per-chunk latency is what bounds the dot warps
      for (int rr = 0; rr < RPS; ++rr)
#pragma unroll
        for (int u = 0; u < RING_U; ++u) accu[rr][u] = 0.0;
#pragma unroll
      for (int rr = 0; rr < RPS; ++rr) {
        if (RPS > 1 && n + rr >= my_rows) break;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          if (c < nch) {
            mbar_wait(&full[s], ph);
            const double2* ch = ring + (size_t)s * RING_CH_PAIRS;
            const int cp = my_pairs - c * RING_CH_PAIRS;         // valid pairs (may exceed a chunk)
#pragma unroll
            for (int u = 0; u < RING_U; ++u) {
              const int p = u * RING_GROUP + tid;
              if (p < cp) {                                      // stale shared memory past the copy
                const double2 x = ch[p];
                accu[rr][u] += x.x * vreg[c][u].x + x.y * vreg[c][u].y;
              }
            }
            if (++s == RING_SLOTS) { s = 0; ph ^= 1u; }
          }
        }
      }
      if (tid == 0) RING_STAMP(1, n);
      if (lane == 0) RING_STAMP2(warp, n);
      double acc[RPS];
#pragma unroll
      for (int rr = 0; rr < RPS; ++rr) acc[rr] = (accu[rr][0] + accu[rr][1]) + (accu[rr][2] + accu[rr][3]);
      if (RPS == 1) acc[0] = warp_sum(acc[0]);
      else warp_sum_k<RPS>(acc);             // (bit-identical to one butterfly per value)
      if (lane == 0) {
#pragma unroll
        for (int rr = 0; rr < RPS; ++rr) dpart[sl][rr][warp] = acc[rr];
        mbar_arrive_local(&dbar[sl]);
      }
      if (++sl == RING_NR) { sl = 0; }
    }
  } else if (warp == 2 * RING_WARPS + 1) {
    // ---------------------------------------------------------------- exchange warp
    // adds the 8 warp partials of a row and posts the CTA's part to every CTA of the cluster.
    // (When the dot warps took turns at this, each turn cost a warp ~3500 cycles of waiting for
    // the other seven, and it needed seven rows to catch up: the slowest dot warp was always
    // ~3000 cycles behind, and so was every row's residual.)
    const unsigned xs_base = smem_u32(&xslot[0][0]);
    const unsigned rd_base = smem_u32(&ready[0]);
    int sl = 0;
    unsigned rph = 0;
    for (long long n = 0; n < my_rows; n += RPS) {
      const int nr = (RPS > 1 && n + RPS > my_rows) ? (int)(my_rows - n) : RPS;   // rows of this step
      double bi[RPS];
#pragma unroll
      for (int rr = 0; rr < RPS; ++rr)       // rank 0 folds -b_i into its part
        bi[rr] = (crank == 0 && rr < nr) ? __ldg(b + i0 + n + rr) : 0.0;
      if (lane == 0) mbar_expect_tx(&ready[sl], 8u * (unsigned)(csize * nr));
      mbar_wait(&dbar[sl], rph);
      if (lane == 0) RING_STAMP(6, n);
      if (lane < csize) {
        const unsigned rb = mapa_u32(rd_base + (unsigned)(sl * 8), (unsigned)lane);
#pragma unroll
        for (int rr = 0; rr < RPS; ++rr) {
          if (rr < nr) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < RING_WARPS; ++w) t += dpart[sl][rr][w];
            t -= bi[rr];
            const unsigned ra = mapa_u32(
                xs_base + (unsigned)((((sl * RPS + rr) * RING_MAX_CLUSTER) + crank) * 8), (unsigned)lane);
            st_async_f64(ra, t, rb);
          }
        }
        if (lane == 0) RING_STAMP(2, n);
      }
      if (++sl == RING_NR) { sl = 0; rph ^= 1u; }
    }
  } else if (warp < 2 * RING_WARPS) {
    // ---------------------------------------------------------------- update warps
    const int ut = tid - RING_GROUP;
    double2 q[NCH][RING_U];
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
      for (int u = 0; u < RING_U; ++u) q[c][u] = make_double2(0.0, 0.0);
    double ss = 0.0;
    int s = 0, sl = 0;
    unsigned ph = 0, rph = 0;
    for (long long n = 0; n < my_rows; n += RPS) {
      const int nr = (RPS > 1 && n + RPS > my_rows) ? (int)(my_rows - n) : RPS;
      double r[RPS];
      {
        mbar_wait(&ready[sl], rph);
        if (ut == 0) RING_STAMP(3, n);
#pragma unroll
        for (int rr = 0; rr < RPS; ++rr) {
          r[rr] = 0.0;
          if (rr < nr)
            for (int c = 0; c < csize; ++c) r[rr] += xslot[sl][rr][c];
          ss += r[rr] * r[rr];
        }
      }
#pragma unroll
      for (int rr = 0; rr < RPS; ++rr) {
        if (RPS > 1 && rr >= nr) break;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          if (c < nch) {
            mbar_wait(&full[s], ph);                             // complete long ago: visibility
            const double2* ch = ring + (size_t)s * RING_CH_PAIRS;
            const int cp = my_pairs - c * RING_CH_PAIRS;
#pragma unroll
            for (int u = 0; u < RING_U; ++u) {
              const int p = u * RING_GROUP + ut;
              if (p < cp) {
                const double2 x = ch[p];
                q[c][u].x += r[rr] * x.x;
                q[c][u].y += r[rr] * x.y;
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_local(&empty[s]);
            if (ut == 0 && c == nch - 1) RING_STAMP(4, n);
            if (++s == RING_SLOTS) { s = 0; ph ^= 1u; }
          }
        }
      }
      if (++sl == RING_NR) { sl = 0; rph ^= 1u; }
    }
    double2* out = reinterpret_cast<double2*>(gpart + (part_base + cid) * n_cols) + p_lo;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
      for (int u = 0; u < RING_U; ++u) {
        const int p = c * RING_CH_PAIRS + u * RING_GROUP + ut;
        if (p < my_pairs) out[p] = q[c][u];
      }
    if (ut == 0 && crank == 0) sq_part[part_base + cid] = ss;
  }
  cluster.sync();          // no CTA leaves while a peer may still post into its shared memory
}

// partial[j] = sum_rb gpart[rb][j] (fixed order);  partial[n_cols] = sum_blk sq_part[blk]
__global__ void __launch_bounds__(256)
lasso_collect_kernel(const double* __restrict__ gpart, int n_rowblocks,
                     const double* __restrict__ sq_part, int n_sq, long long n_cols,
                     double* __restrict__ partial) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gpart && j < n_cols) {
    double s = 0.0;
    for (int rb = 0; rb < n_rowblocks; ++rb) s += gpart[(long long)rb * n_cols + j];
    partial[j] = s;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double s = 0.0;
    for (int k = 0; k < n_sq; ++k) s += sq_part[k];
    partial[n_cols] = s;
  }
}

constexpr int VEC_THREADS = 256;
constexpr int VEC_MAX_BLOCKS = 1024;

// block partials: sums[0..2] added, sums[3] max-ed; final combine over blocks in index order
struct StepSums { double gd, dd, abs1, maxd; };

__device__ __forceinline__ void block_reduce_step(StepSums& s, StepSums* sh) {
  s.gd = warp_sum(s.gd);
  s.dd = warp_sum(s.dd);
  s.abs1 = warp_sum(s.abs1);
  s.maxd = warp_max(s.maxd);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sh[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    StepSums t = sh[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      t.gd += sh[w].gd; t.dd += sh[w].dd; t.abs1 += sh[w].abs1; t.maxd = fmax(t.maxd, sh[w].maxd);
    }
    s = t;
  }
}

// x = soft(y - lr*g, lr*l1), g = partial*(2*scale)  (test_proximal_gradient.py:55-63), plus
//   gd = g.(x-y), dd = ||x-y||^2 (sum), abs1 = ||x||_1, maxd = max|x-y|
// If y == nullptr only abs1 of x_in is produced (g(x0) at start-up).
__global__ void __launch_bounds__(VEC_THREADS)
lasso_prox_kernel(const double* __restrict__ y, const double* __restrict__ partial,
                  double two_scale, double lr, double thresh, long long n,
                  double* __restrict__ x, double* __restrict__ g_out,
                  StepSums* __restrict__ block_sums, unsigned int* __restrict__ counter,
                  StepSums* __restrict__ out) {
  __shared__ StepSums sh[VEC_THREADS / 32];
  __shared__ bool is_last;
  StepSums s{0.0, 0.0, 0.0, 0.0};
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n;
       j += (long long)gridDim.x * blockDim.x) {
    if (y) {
      const double gj = partial[j] * two_scale;
      const double yj = y[j];
      const double xj = soft_threshold(fma(-lr, gj, yj), thresh);
      const double d = xj - yj;
      x[j] = xj;
      if (g_out) g_out[j] = gj;
      s.gd = fma(gj, d, s.gd);
      s.dd = fma(d, d, s.dd);
      s.abs1 += fabs(xj);
      s.maxd = fmax(s.maxd, fabs(d));
    } else {
      s.abs1 += fabs(x[j]);
    }
  }
  block_reduce_step(s, sh);
  if (threadIdx.x == 0) {
    block_sums[blockIdx.x] = s;
    __threadfence();
    const unsigned int done = atomicAdd(counter, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    StepSums t = block_sums[0];
    for (unsigned int k = 1; k < gridDim.x; ++k) {
      const StepSums u = block_sums[k];
      t.gd += u.gd; t.dd += u.dd; t.abs1 += u.abs1; t.maxd = fmax(t.maxd, u.maxd);
    }
    *out = t;
    *counter = 0u;
  }
}

// y = x + mom*(x - x_prev)   (proximal_gradient.py:534)
__global__ void __launch_bounds__(VEC_THREADS)
lasso_momentum_kernel(const double* __restrict__ x, const double* __restrict__ xp, double mom,
                      long long n, double* __restrict__ y) {
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n;
       j += (long long)gridDim.x * blockDim.x) {
    const double xj = x[j];
    y[j] = fma(mom, xj - xp[j], xj);
  }
}

// grad = partial * (2*scale) (bench / zf_lasso_gradient_device)
__global__ void __launch_bounds__(VEC_THREADS)
lasso_scale_kernel(const double* __restrict__ partial, double two_scale, double scale,
                   long long n, double* __restrict__ grad, double* __restrict__ f_out) {
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n;
       j += (long long)gridDim.x * blockDim.x)
    grad[j] = partial[j] * two_scale;
  if (f_out && blockIdx.x == 0 && threadIdx.x == 0) {
    const double nrm = sqrt(partial[n]);
    *f_out = nrm * nrm * scale;
  }
}


// ---------------------------------------------------------------------------------------
// Device-decided loop (zf_lasso_dev_*): the scalars of proximal_gradient.py:474-538 -- lr, t_k,
// F values, the line-search and stop decisions -- live in ONE device struct, every kernel of an
// iteration reads what it needs from it, and the last block of the n_cols-sized kernel (or a
// one-thread kernel) takes the decision the host used to take.  Nothing synchronises with the
// host inside the loop: the host enqueues "slots" (one trial each) ahead of the GPU, as a CUDA
// graph when it can, and polls the state every few dozen slots.  Once the state says DONE the
// remaining slots are no-ops (every kernel tests a flag first).
//
// The scalar arithmetic uses the round-to-nearest intrinsics so that nvcc cannot contract it
// into FMAs: it is bit for bit the host code of zf_lasso_step / accept_candidate above.
// ---------------------------------------------------------------------------------------
enum { DV_GRAD = 0, DV_RETRY = 1, DV_DONE = 2 };

struct LassoDevOpts {
  double lr0, tol, tol_internal, decay, na, nb, scale, l1;
  long long max_iter;
  int max_bt, nesterov, deprecated, need_F, cap;
  double* allerrs;        // device traces (cap, cap + 1) or nullptr
  double* allfuns;
  double* allvecs;        // (vec_rows x n_cols) iterates x^0 .. or nullptr (zf_lasso_set_allvecs)
  long long vec_rows;
};

struct LassoDevState {
  double lr, t_prev, F_prev, F_x, f_y, sub_fun, err, mom;
  StepSums sums;          // of the candidate now in x_new
  long long nit;
  int status, phase, bt, accept;
  long long record_nit;   // > 0: the momentum kernel stores x_new as iterate `record_nit` (allvecs)
  int skip_grad;          // != 0: the gradient pass of this slot has nothing to do (retry / done)
  int done;               // != 0: the solve has ended
  int F_known, result_is_prev;
  unsigned int ticket;    // last-block election of the update kernel
  int p2p_error;          // a peer's flag did not arrive in time (row-sharded peer exchange)
  unsigned long long gseq, sseq;   // gradient / residual-norm exchanges published so far
};

// ---------------------------------------------------------------------------------------
// Row-sharded runs on one node: the exchange of [A^T r | sum r^2] folded into the kernels.
// Every rank owns one P2PBuf in its own HBM, mapped into every other rank of the node
// (cudaIpc handles, NVLink peer access).  A rank PUBLISHES its collected partials into its own
// buffer and then bumps the buffer's flag; the consumers -- the same update / decide kernels
// that a single GPU runs -- wait for every peer's flag and add the peers' partials straight out
// of peer memory, in rank order, so every rank forms bit-identical sums.  No NCCL launch, no
// extra pass: the all-reduce IS the operand fetch of the prox kernel.  Buffers alternate with the
// exchange's sequence number; a rank reaches exchange k + 2 (which overwrites buffer k & 1) only
// after its own exchange k + 1 saw every peer's flag k + 1, i.e. after every peer finished
// reading exchange k.  Flags are monotonic, waits are bounded (an error flag, never a hang).
// ---------------------------------------------------------------------------------------
constexpr int P2P_MAX_RANKS = 8;
struct P2PBuf {
  unsigned long long flag_g[2];      // sequence number of the gradient partials in data[p]
  unsigned long long flag_s[2];      // sequence number of ss[p]
  double ss[2];
  double pad[2];
  double data[1];                    // [2][stride]: A^T r partial (n_cols) | sum r^2
};
struct P2PPeers {
  P2PBuf* buf[P2P_MAX_RANKS];        // every rank's buffer, own included, in rank order
  long long stride;                  // doubles per data slot (>= n_cols + 1)
  int world, rank;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
// wait until *flag >= seq (one thread); false after ~10 s (2^34 cycles): ranks enter a solve
// within milliseconds of each other (the wrapper puts a barrier in front of it), a peer that
// is this late has died
__device__ bool p2p_wait(const unsigned long long* flag, unsigned long long seq) {
  const long long t0 = clock64();
  while (ld_acquire_sys_u64(flag) < seq) {
    __nanosleep(64);
    if (clock64() - t0 > (1LL << 34)) return false;
  }
  return true;
}

// publish the collected gradient partials: mine->data[p][j] = sum_rb gpart[rb][j], [n_cols] = ss
__global__ void __launch_bounds__(256)
lasso_p2p_publish_kernel(LassoDevState* st, const double* __restrict__ gpart, int n_parts,
                         const double* __restrict__ sq_part, int n_sq, long long n_cols,
                         P2PPeers pp, unsigned int* __restrict__ ticket) {
  if (st->done != 0 || st->skip_grad != 0) return;
  const unsigned long long seq = st->gseq + 1ull;
  P2PBuf* mine = pp.buf[pp.rank];
  double* out = mine->data + (seq & 1ull) * pp.stride;
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n_cols) {
    double acc = 0.0;
    for (int rb = 0; rb < n_parts; ++rb) acc += gpart[(long long)rb * n_cols + j];
    out[j] = acc;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double acc = 0.0;
    for (int k = 0; k < n_sq; ++k) acc = __dadd_rn(acc, sq_part[k]);
    out[n_cols] = acc;
  }
  __shared__ bool is_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    *ticket = 0u;
    __threadfence_system();
    st_release_sys_u64(&mine->flag_g[seq & 1ull], seq);
    st->gseq = seq;
  }
}

// publish this rank's sum r^2 (residual pass at the candidate / x0 / the final x)
__global__ void lasso_p2p_publish_ss_kernel(LassoDevState* st, const double* __restrict__ sq_part,
                                            int n_sq, P2PPeers pp, int force) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (st->done != 0 && !force) return;
  const unsigned long long seq = st->sseq + 1ull;
  P2PBuf* mine = pp.buf[pp.rank];
  double acc = 0.0;
  for (int k = 0; k < n_sq; ++k) acc = __dadd_rn(acc, sq_part[k]);
  mine->ss[seq & 1ull] = acc;
  __threadfence_system();
  st_release_sys_u64(&mine->flag_s[seq & 1ull], seq);
  st->sseq = seq;
}

// every rank's sum r^2 of the exchange just published, in rank order, into local memory
__global__ void lasso_p2p_gather_ss_kernel(LassoDevState* st, P2PPeers pp, double* __restrict__ out,
                                           int force) {
  if (blockIdx.x != 0 || (int)threadIdx.x >= pp.world) return;
  if (st->done != 0 && !force) return;
  const unsigned long long seq = st->sseq;
  const P2PBuf* peer = pp.buf[threadIdx.x];
  if (!p2p_wait(&peer->flag_s[seq & 1ull], seq)) {
    atomicExch(&st->p2p_error, 1);
    out[threadIdx.x] = 0.0;
    return;
  }
  out[threadIdx.x] = ld_relaxed_sys_f64(&peer->ss[seq & 1ull]);
}

__device__ __forceinline__ double dv_f_from_ss(double ss, double scale) {
  const double nrm = __dsqrt_rn(ss);
  return __dmul_rn(__dmul_rn(nrm, nrm), scale);
}

__device__ __forceinline__ void dv_next_momentum(const LassoDevOpts& o, double t, double* t_new,
                                                 double* mom) {
  // t_{k+1} = sqrt(t_k^2 - a t_k + b) + 1/2 ; mom = (t_k - 1) / t_{k+1}   (proximal_gradient.py:531-534)
  if (o.nesterov) {
    const double tn = __dadd_rn(
        __dsqrt_rn(__dadd_rn(__dsub_rn(__dmul_rn(t, t), __dmul_rn(o.na, t)), o.nb)), 0.5);
    *t_new = tn;
    *mom = __ddiv_rn(__dsub_rn(t, 1.0), tn);
  } else {
    *t_new = t;
    *mom = 0.0;
  }
}

// The candidate in x_new was accepted (one thread): stop test, trace, t_{k+1}; returns true if the
// loop goes on (the caller then applies the momentum / rotation).
__device__ bool dv_accept(const LassoDevOpts& o, LassoDevState* st, double maxd) {
  st->err = maxd;
  const long long nit = st->nit;
  st->record_nit = (o.allvecs && nit < o.vec_rows) ? nit : 0;
  if (o.cap > 0 && nit <= o.cap) {
    if (o.allerrs) o.allerrs[nit - 1] = maxd;
    if (o.allfuns && st->F_known) o.allfuns[nit] = st->F_x;
  }
  const bool converged = maxd < o.tol;
  if (converged || nit >= o.max_iter) {
    st->status = converged ? 1 : 0;
    st->result_is_prev = 0;
    st->phase = DV_DONE;
    st->done = 1;
    st->skip_grad = 1;
    st->accept = 0;
    return false;
  }
  double t_new, mom;
  dv_next_momentum(o, st->t_prev, &t_new, &mom);
  st->t_prev = t_new;
  st->mom = mom;
  st->F_prev = st->F_x;
  st->nit = nit + 1;
  st->phase = DV_GRAD;
  st->skip_grad = 0;
  st->bt = 0;
  return true;
}

// F(x0) from the residual norm of x0 (ss_src[0]) and ||x0||_1 (sums->abs1)
__global__ void lasso_dev_init_kernel(const LassoDevOpts* __restrict__ op, LassoDevState* st,
                                      const StepSums* __restrict__ sums,
                                      const double* __restrict__ ss_src, int n_sq) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const LassoDevOpts o = *op;
  double ss = 0.0;
  for (int k = 0; k < n_sq; ++k) ss = __dadd_rn(ss, ss_src[k]);
  const double F0 = __dadd_rn(dv_f_from_ss(ss, o.scale), __dmul_rn(o.l1, sums->abs1));
  st->lr = o.lr0;
  st->t_prev = 1.0;
  st->F_prev = F0;
  st->F_x = F0;
  st->f_y = 0.0;
  st->sub_fun = 0.0;
  st->err = CUDART_INF;
  st->mom = 0.0;
  st->nit = 1;
  st->status = 0;
  st->phase = DV_GRAD;
  st->bt = 0;
  st->accept = 0;
  st->skip_grad = 0;
  st->done = 0;
  st->F_known = 0;
  st->result_is_prev = 0;
  st->record_nit = 0;
  st->ticket = 0u;
  // (gseq / sseq run on across solves; p2p_error is sticky: the F(x0) exchange precedes this kernel)
  if (st->p2p_error) {
    st->status = -3;
    st->phase = DV_DONE;
    st->done = 1;
    st->skip_grad = 1;
  }
  if (o.cap > 0 && o.allfuns) o.allfuns[0] = F0;
}

// One trial:  x_new = soft(y - lr g, lr l1)  with  g = 2 scale A^T(A y - b)  taken from
//   SRC 0: the row-block partials gpart (the collect pass is folded in; sum r^2 from sq_part),
//   SRC 1: `partial` (already collected and, in a row-sharded run, all-reduced),
//   a retry (state.phase == DV_RETRY): the g stored by the trial that computed it.
// FIXED (decay_rate == 1, no F trace): the step is accepted unconditionally, so the same sweep
// also writes y^{k+1} = x + mom (x - x_prev) and x_prev = x, and the last block runs the stop test
// and t_{k+1}: a whole iteration is the gradient pass plus this kernel.
// Otherwise the last block leaves f(y) and the subproblem value for lasso_dev_decide_kernel.
template <int SRC, bool FIXED>
__global__ void __launch_bounds__(VEC_THREADS)
lasso_dev_update_kernel(const LassoDevOpts* __restrict__ op, LassoDevState* st,
                        const double* __restrict__ gpart, int n_parts,
                        const double* __restrict__ sq_part, int n_sq,
                        const double* __restrict__ partial, long long n,
                        double* __restrict__ y, double* __restrict__ xp, double* __restrict__ xn,
                        double* __restrict__ g, StepSums* __restrict__ block_sums, P2PPeers pp) {
  if (*reinterpret_cast<const volatile int*>(&st->done) != 0) return;
  __shared__ StepSums sh[VEC_THREADS / 32];
  __shared__ bool is_last;
  const LassoDevOpts o = *op;
  const double lr = st->lr;
  const bool retry = (st->phase == DV_RETRY);
  // SRC 2: the gradient is the sum of every rank's published partials, read from peer memory
  const unsigned long long gseq = (SRC == 2) ? st->gseq : 0ull;
  const double* peer_data[P2P_MAX_RANKS];
  if (SRC == 2 && !retry) {
    if ((int)threadIdx.x < pp.world) {
      if (!p2p_wait(&pp.buf[threadIdx.x]->flag_g[gseq & 1ull], gseq)) atomicExch(&st->p2p_error, 1);
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < P2P_MAX_RANKS; ++r)
      peer_data[r] = (r < pp.world) ? pp.buf[r]->data + (gseq & 1ull) * pp.stride : nullptr;
  }
  const double two_scale = 2.0 * o.scale;
  // the rounded product, as the host-decided loop passes it and as numpy forms l1 * lr: nvcc must
  // not contract it into the subtraction inside soft_threshold
  const double thresh = __dmul_rn(o.l1, lr);
  double t_new = 0.0, mom = 0.0;
  if (FIXED) dv_next_momentum(o, st->t_prev, &t_new, &mom);
  StepSums s{0.0, 0.0, 0.0, 0.0};
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n;
       j += (long long)gridDim.x * blockDim.x) {
    double gj;
    if (retry) {
      gj = g[j];
    } else if (SRC == 0) {
      double acc = 0.0;
      for (int rb = 0; rb < n_parts; ++rb) acc += gpart[(long long)rb * n + j];
      gj = acc * two_scale;
    } else if (SRC == 2) {
      double acc = 0.0;
#pragma unroll
      for (int r = 0; r < P2P_MAX_RANKS; ++r)
        if (r < pp.world) acc += ld_relaxed_sys_f64(peer_data[r] + j);
      gj = acc * two_scale;
    } else {
      gj = partial[j] * two_scale;
    }
    const double yj = y[j];
    const double xj = soft_threshold(fma(-lr, gj, yj), thresh);
    const double d = xj - yj;
    xn[j] = xj;
    s.gd = fma(gj, d, s.gd);
    s.dd = fma(d, d, s.dd);
    s.abs1 += fabs(xj);
    s.maxd = fmax(s.maxd, fabs(d));
    if (FIXED) {
      y[j] = fma(mom, xj - xp[j], xj);
      xp[j] = xj;
    } else if (!retry) {
      g[j] = gj;
    }
  }
  block_reduce_step(s, sh);
  if (threadIdx.x == 0) {
    block_sums[blockIdx.x] = s;
    __threadfence();
    const unsigned int done = atomicAdd(&st->ticket, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last || threadIdx.x != 0) return;
  __threadfence();
  StepSums t = block_sums[0];
  for (unsigned int k = 1; k < gridDim.x; ++k) {
    const StepSums u = block_sums[k];
    t.gd += u.gd; t.dd += u.dd; t.abs1 += u.abs1; t.maxd = fmax(t.maxd, u.maxd);
  }
  st->ticket = 0u;
  st->sums = t;
  st->accept = 0;
  if (SRC == 2 && st->p2p_error) {      // a peer never published: stop with an error status
    st->status = -3;
    st->phase = DV_DONE;
    st->done = 1;
    st->skip_grad = 1;
    return;
  }
  if (FIXED) {
    st->F_known = 0;
    dv_accept(o, st, t.maxd);
  } else {
    if (!retry) {
      double ss = 0.0;
      if (SRC == 0) {
        for (int k = 0; k < n_sq; ++k) ss = __dadd_rn(ss, sq_part[k]);
      } else if (SRC == 2) {
        for (int r = 0; r < pp.world; ++r) ss = __dadd_rn(ss, ld_relaxed_sys_f64(peer_data[r] + n));
      } else {
        ss = partial[n];
      }
      st->f_y = dv_f_from_ss(ss, o.scale);
    }
    // proximal_gradient.py:149-155 for one objective
    const double nrm = __dsqrt_rn(t.dd);
    double fun = __dadd_rn(__dadd_rn(t.gd, __dmul_rn(o.l1, t.abs1)),
                           __ddiv_rn(__ddiv_rn(__dmul_rn(nrm, nrm), 2.0), lr));
    if (!o.deprecated) fun = __dadd_rn(fun, __dsub_rn(st->f_y, st->F_prev));
    st->sub_fun = fun;
  }
}

// After the residual pass at the candidate: F(x_new), the line-search test, then either the
// acceptance logic or a smaller step (one thread).  ss_src: sq_part (n_sq values) or
// &partial[n_cols] (1 value, all-reduced).
__global__ void lasso_dev_decide_kernel(const LassoDevOpts* __restrict__ op, LassoDevState* st,
                                        const double* __restrict__ ss_src, int n_sq) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (st->done) return;
  if (st->p2p_error) {                 // a peer's residual norm never arrived
    st->status = -3;
    st->phase = DV_DONE;
    st->done = 1;
    st->skip_grad = 1;
    st->accept = 0;
    return;
  }
  const LassoDevOpts o = *op;
  st->record_nit = 0;                  // set again by dv_accept if this candidate is accepted
  double ss = 0.0;
  for (int k = 0; k < n_sq; ++k) ss = __dadd_rn(ss, ss_src[k]);
  const double f_x = dv_f_from_ss(ss, o.scale);
  st->F_x = __dadd_rn(f_x, __dmul_rn(o.l1, st->sums.abs1));
  st->F_known = 1;
  bool ok;
  if (o.decay == 1.0) ok = true;                                            // proximal_gradient.py:298
  else if (o.deprecated) ok = (__dsub_rn(f_x, st->f_y) <= __dadd_rn(st->sub_fun, o.tol_internal));
  else ok = (__dsub_rn(st->F_x, st->F_prev) <= __dadd_rn(st->sub_fun, o.tol_internal));
  if (ok) {
    st->accept = dv_accept(o, st, st->sums.maxd) ? 1 : 0;
    return;
  }
  st->lr = __dmul_rn(st->lr, o.decay);
  st->bt += 1;
  st->accept = 0;
  if (st->bt >= o.max_bt) {
    // RuntimeError("Backtracking failed ...") -> x = x_prev, nit - 1 (proximal_gradient.py:493-509)
    st->result_is_prev = 1;
    st->F_x = st->F_prev;
    st->nit -= 1;
    st->status = -1;
    st->phase = DV_DONE;
    st->done = 1;
    st->skip_grad = 1;
  } else {
    st->phase = DV_RETRY;     // same gradient, smaller step
    st->skip_grad = 1;
  }
}

// y = x + mom (x - x_prev), x_prev = x  when the decide kernel accepted the candidate
__global__ void __launch_bounds__(VEC_THREADS)
lasso_dev_momentum_kernel(const LassoDevOpts* __restrict__ op,
                          const LassoDevState* __restrict__ st, long long n,
                          const double* __restrict__ xn, double* __restrict__ xp,
                          double* __restrict__ y) {
  // return_all's allvecs: the accepted candidate is iterate `record_nit` (also when the solve
  // ends with it; slots after the end rewrite the same row with the same values)
  const long long rec = st->record_nit;
  if (rec > 0) {
    double* row = op->allvecs + rec * n;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n;
         j += (long long)gridDim.x * blockDim.x)
      row[j] = xn[j];
  }
  if (st->accept == 0 || st->done != 0) return;
  const double mom = st->mom;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n;
       j += (long long)gridDim.x * blockDim.x) {
    const double xj = xn[j];
    y[j] = fma(mom, xj - xp[j], xj);
    xp[j] = xj;
  }
}

// res.fun when the loop ran without F evaluations: F(x) from one last residual pass
__global__ void lasso_dev_final_kernel(const LassoDevOpts* __restrict__ op, LassoDevState* st,
                                       const double* __restrict__ ss_src, int n_sq) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (st->F_known || st->result_is_prev) return;
  const LassoDevOpts o = *op;
  double ss = 0.0;
  for (int k = 0; k < n_sq; ++k) ss = __dadd_rn(ss, ss_src[k]);
  st->F_x = __dadd_rn(dv_f_from_ss(ss, o.scale), __dmul_rn(o.l1, st->sums.abs1));
  st->F_known = 1;
}

}  // namespace zf

// =======================================================================================
// handle
// =======================================================================================
enum LassoPhase { LP_IDLE = 0, LP_INIT, LP_GRAD, LP_FNEW, LP_FINAL, LP_DONE };

struct zf_lasso {
  const double* A = nullptr;
  const double* b = nullptr;
  long long n_rows = 0, n_cols = 0;
  double scale = 1.0, l1 = 0.0;
  cudaStream_t st = 0;
  bool vec = false;
  int n_sm = 148;
  // device workspace
  double* vecs = nullptr;       // 4 * n_cols: x_prev, x_new, y, g
  double *xp = nullptr, *xn = nullptr, *y = nullptr, *g = nullptr;
  double* r = nullptr;          // n_rows
  double* gpart = nullptr;      // n_rowblocks * n_cols
  double* sq_part = nullptr;    // res_blocks
  double* partial = nullptr;    // n_cols + 1
  zf::StepSums* block_sums = nullptr;
  zf::StepSums* d_sums = nullptr;
  unsigned int* counter = nullptr;
  double* h_pin = nullptr;      // pinned: [StepSums (4) | ss]
  int res_blocks = 0, n_slabs = 0, n_rowblocks = 0, vec_blocks = 0;
  long long rows_per_block = 0;
  // fused one-pass gradient (0 = not applicable)
  int fused_pairs = 0, fused_ctas = 0;
  long long fused_rows_per_cta = 0;
  int fused_cluster = 1;              // CTAs per cluster (1: single-CTA kernel)
  bool fused_ring = false;            // warp-specialised chunk ring (lasso_fused_ring_kernel)
  // second, concurrent ring launch on the SMs the first one cannot use (4-CTA clusters fit on
  // 132 of 148 SMs): 2-CTA clusters over the last rows, on its own stream
  int ring2_ctas = 0, ring2_nch = 0, ring2_cluster = 2, ring1_clusters = 0;
  long long ring2_row0 = 0, ring2_rows_per_cluster = 0, ring2_pairs_per_cta = 0;
  long long ring_rows = 0;            // rows the ring plan covers (n_rows; a prefix while tuning)
  int tuned_cluster = 0;              // what the create-time probe chose (0: static policy)
  double tuned_rate2 = 0.0;
  cudaStream_t st2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  long long fused_pairs_per_cta = 0;  // column pairs per CTA (cluster form)
  size_t gpart_rows = 0;
  // solver state (host scalars)
  zf_options opt{};
  int phase = LP_IDLE;
  double lr = 1.0, t_prev = 1.0, F_prev = 0.0, F_x = 0.0, f_y = 0.0, sub_fun = 0.0, err = 0.0;
  bool F_known = false, need_F = false;
  long long nit = 0;
  int status = 0, bt = 0;
  zf::StepSums sums{};
  bool result_is_prev = false;
  const int* skip = nullptr;    // device flag the gradient-pass kernels test (device-decided loop)
  int ring_rps = ZF_RING_RPS_DEFAULT;   // rows per exchange step of the chunk-ring kernel (create())
  // device-decided loop (zf_lasso_dev_*)
  zf::LassoDevOpts* d_opts = nullptr;
  zf::LassoDevState* d_state = nullptr;
  zf::LassoDevState* h_state = nullptr;      // pinned, 2 poll slots
  cudaEvent_t ev_poll[2] = {nullptr, nullptr};
  double* d_allerrs = nullptr;
  double* d_allfuns = nullptr;
  int dev_cap = 0;
  double* h_allvecs = nullptr;               // zf_lasso_set_allvecs: host target of the next solve
  long long h_allvecs_rows = 0;
  double* d_allvecs = nullptr;
  long long d_allvecs_rows = 0;
  bool dev_sharded = false, dev_fixed = false, dev_active = false;
  zf::LassoDevOpts dev_opts_host{};
  cudaStream_t st_own = nullptr;             // used when the caller's stream cannot be captured
  cudaEvent_t ev_own = nullptr;
  // row-sharded peer exchange (zf_lasso_p2p_*)
  zf::P2PPeers pp{};
  zf::P2PBuf* p2p_mine = nullptr;
  void* p2p_opened[zf::P2P_MAX_RANKS] = {};
  bool p2p_on = false;
  double* ss_gather = nullptr;               // every rank's sum r^2, in rank order
  unsigned int* p2p_ticket = nullptr;
  cudaGraphExec_t graph[2] = {nullptr, nullptr};   // [fixed-step slots, line-search slots]
  int graph_slots = 0;
  bool graph_failed = false;
  double* h_allerrs = nullptr;
  double* h_allfuns = nullptr;
};

namespace {

#define ZF_CUDA(call)                                          \
  do {                                                         \
    cudaError_t _e = (call);                                   \
    if (_e != cudaSuccess) return zf::zf_fail_cuda(_e, #call); \
  } while (0)

int launch_residual(zf_lasso* h, const double* v) {
  if (h->vec)
    zf::lasso_residual_kernel<true><<<h->res_blocks, zf::RES_THREADS, 0, h->st>>>(
        h->A, h->b, v, h->n_rows, h->n_cols, h->r, h->sq_part, h->skip);
  else
    zf::lasso_residual_kernel<false><<<h->res_blocks, zf::RES_THREADS, 0, h->st>>>(
        h->A, h->b, v, h->n_rows, h->n_cols, h->r, h->sq_part, h->skip);
  ZF_CUDA(cudaGetLastError());
  zf::zf_count_launch();
  return ZF_OK;
}

int launch_atr(zf_lasso* h) {
  dim3 grid((unsigned)h->n_slabs, (unsigned)h->n_rowblocks);
  if (h->vec)
    zf::lasso_atr_kernel<true><<<grid, zf::ATR_THREADS, 0, h->st>>>(
        h->A, h->r, h->n_rows, h->n_cols, h->rows_per_block, h->gpart, h->skip);
  else
    zf::lasso_atr_kernel<false><<<grid, zf::ATR_THREADS, 0, h->st>>>(
        h->A, h->r, h->n_rows, h->n_cols, h->rows_per_block, h->gpart, h->skip);
  ZF_CUDA(cudaGetLastError());
  zf::zf_count_launch();
  return ZF_OK;
}

template <int PAIRS>
int launch_fused_t(zf_lasso* h, const double* v) {
  auto k = zf::lasso_fused_kernel<PAIRS>;
  const size_t smem = sizeof(double) * (size_t)h->n_cols;
  ZF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<h->fused_ctas, zf::FUSED_THREADS, smem, h->st>>>(h->A, h->b, v, h->n_rows, h->n_cols,
                                                       h->fused_rows_per_cta, h->gpart, h->sq_part,
                                                       h->skip);
  ZF_CUDA(cudaGetLastError());
  zf::zf_count_launch();
  return ZF_OK;
}

struct RingLaunch {
  int ctas, cluster;
  long long rows_per_cluster, pairs_per_cta, row_begin, row_end;
  int part_base;
  cudaStream_t stream;
};

template <int NCH, int RPS>
int launch_fused_ring_t(zf_lasso* h, const double* v, const RingLaunch& L, bool query_only,
                        int* max_clusters) {
  auto k = zf::lasso_fused_ring_kernel<NCH, RPS>;
  const size_t smem = (size_t)zf::RING_SLOTS * zf::RING_CH_PAIRS * 16;
  ZF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)L.ctas);
  cfg.blockDim = dim3(zf::RING_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = L.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)L.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (query_only) {
    ZF_CUDA(cudaOccupancyMaxActiveClusters(max_clusters, k, &cfg));
    return ZF_OK;
  }
  // ZF_LASSO_RING_DBG=<block index>: which CTA of the grid writes its stamps
  static const int dbg = (ZF_RING_DEBUG && getenv("ZF_LASSO_RING_DBG"))
                             ? atoi(getenv("ZF_LASSO_RING_DBG")) + 1 : 0;
  ZF_CUDA(cudaLaunchKernelEx(&cfg, k, h->A, h->b, v, h->n_rows, h->n_cols, L.rows_per_cluster,
                             L.pairs_per_cta, h->gpart, h->sq_part, dbg, L.row_begin,
                             L.row_end, L.part_base, h->skip));
  zf::zf_count_launch();
  if (dbg && L.part_base == 0) {
    long long st[8][64], st2[8][64];
    ZF_CUDA(cudaStreamSynchronize(L.stream));
    ZF_CUDA(cudaMemcpyFromSymbol(st, zf::zf_ring_dbg, sizeof(st)));
    ZF_CUDA(cudaMemcpyFromSymbol(st2, zf::zf_ring_dbg2, sizeof(st2)));
    for (int r = 16; r < 24; ++r) {
      fprintf(stderr, "[ring dbg2] row %3d dot done per warp:", 64 + r);
      for (int w = 0; w < 8; ++w) fprintf(stderr, " %6lld", st2[w][r] - st[0][r]);
      fprintf(stderr, "\n");
    }
    for (int r = 0; r < 64; r += 4)
      fprintf(stderr, "[ring dbg] row %3d: issue 0  warp0 dot done %6lld  all warps in %6lld  posted %6lld  ready %6lld  released %6lld | next issue %6lld\n",
              64 + r, st[1][r] - st[0][r], st[6][r] - st[0][r], st[2][r] - st[0][r],
              st[3][r] - st[0][r], st[4][r] - st[0][r], st[0][r + 1] - st[0][r]);
  }
  return ZF_OK;
}

int launch_fused_ring_n(zf_lasso* h, const double* v, int nch, const RingLaunch& L, bool query_only,
                        int* max_clusters) {
  // two rows per exchange step for narrow row slices (<= 3 chunks); ZF_LASSO_RING_RPS=1|2 overrides
  // rows per exchange step: the chunks of one step stay in the ring until its update is done, so
  // a step may hold at most 6 of the 13 slots or the producer runs out of prefetch depth
  // (measured: 2 rows x 4 chunks drops from 1.00 to 0.66 of the copy bandwidth; 3 rows x <= 2
  // chunks gains nothing over 2 rows, profiles/r02_ring_rows_per_step.txt)
  if (h->ring_rps == 2 && nch <= 3) {
    switch (nch) {
      case 1: return launch_fused_ring_t<1, 2>(h, v, L, query_only, max_clusters);
      case 2: return launch_fused_ring_t<2, 2>(h, v, L, query_only, max_clusters);
      default: return launch_fused_ring_t<3, 2>(h, v, L, query_only, max_clusters);
    }
  }
  switch (nch) {                     // chunks per row slice
    case 1: return launch_fused_ring_t<1, 1>(h, v, L, query_only, max_clusters);
    case 2: return launch_fused_ring_t<2, 1>(h, v, L, query_only, max_clusters);
    case 3: return launch_fused_ring_t<3, 1>(h, v, L, query_only, max_clusters);
    case 4: return launch_fused_ring_t<4, 1>(h, v, L, query_only, max_clusters);
    default: return launch_fused_ring_t<5, 1>(h, v, L, query_only, max_clusters);
  }
}

int launch_fused_ring(zf_lasso* h, const double* v, bool query_only, int* max_clusters) {
  RingLaunch L1{h->fused_ctas, h->fused_cluster, h->fused_rows_per_cta, h->fused_pairs_per_cta, 0,
                h->ring2_ctas > 0 ? h->ring2_row0 : h->ring_rows, 0, h->st};
  if (query_only || h->ring2_ctas == 0)
    return launch_fused_ring_n(h, v, h->fused_pairs, L1, query_only, max_clusters);
  // fork: the wide launch first (it takes every SM a 4-CTA cluster fits on), then the 2-CTA
  // launch for the SMs it left idle; join on the handle's stream
  ZF_CUDA(cudaEventRecord(h->ev_fork, h->st));
  ZF_CUDA(cudaStreamWaitEvent(h->st2, h->ev_fork, 0));
  int rc = launch_fused_ring_n(h, v, h->fused_pairs, L1, false, nullptr);
  if (rc != ZF_OK) return rc;
  RingLaunch L2{h->ring2_ctas, h->ring2_cluster, h->ring2_rows_per_cluster, h->ring2_pairs_per_cta, h->ring2_row0,
                h->ring_rows, h->ring1_clusters, h->st2};
  rc = launch_fused_ring_n(h, v, h->ring2_nch, L2, false, nullptr);
  if (rc != ZF_OK) return rc;
  ZF_CUDA(cudaEventRecord(h->ev_join, h->st2));
  ZF_CUDA(cudaStreamWaitEvent(h->st, h->ev_join, 0));
  return ZF_OK;
}

// gradient pass at v: leaves the A^T r partials in gpart (n_gpart_rows x n_cols) and the
// sum r^2 partials in sq_part (n_sq), by the fused kernel when it applies
int launch_gradient_pass(zf_lasso* h, const double* v, int* n_gpart_rows, int* n_sq) {
  if (h->fused_pairs > 0 && h->fused_ring) {
    const int rc = launch_fused_ring(h, v, false, nullptr);
    *n_gpart_rows = h->fused_ctas / h->fused_cluster + h->ring2_ctas / h->ring2_cluster;
    *n_sq = *n_gpart_rows;
    return rc;
  }
  if (h->fused_pairs > 0) {
    int rc;
    switch (h->fused_pairs) {
      case 4: rc = launch_fused_t<4>(h, v); break;
      case 8: rc = launch_fused_t<8>(h, v); break;
      case 12: rc = launch_fused_t<12>(h, v); break;
      case 16: rc = launch_fused_t<16>(h, v); break;
      default: rc = launch_fused_t<20>(h, v); break;
    }
    *n_gpart_rows = h->fused_ctas;
    *n_sq = h->fused_ctas;
    return rc;
  }
  int rc = launch_residual(h, v);
  if (rc != ZF_OK) return rc;
  rc = launch_atr(h);
  *n_gpart_rows = h->n_rowblocks;
  *n_sq = h->res_blocks;
  return rc;
}

int launch_collect_n(zf_lasso* h, bool with_gradient, int n_gpart_rows, int n_sq) {
  const int blocks = with_gradient ? (int)((h->n_cols + 255) / 256) : 1;
  zf::lasso_collect_kernel<<<blocks, 256, 0, h->st>>>(with_gradient ? h->gpart : nullptr,
                                                     n_gpart_rows, h->sq_part, n_sq, h->n_cols,
                                                     h->partial);
  ZF_CUDA(cudaGetLastError());
  zf::zf_count_launch();
  return ZF_OK;
}

int launch_collect(zf_lasso* h, bool with_gradient) {
  const int blocks = with_gradient ? (int)((h->n_cols + 255) / 256) : 1;
  zf::lasso_collect_kernel<<<blocks, 256, 0, h->st>>>(with_gradient ? h->gpart : nullptr,
                                                     h->n_rowblocks, h->sq_part, h->res_blocks,
                                                     h->n_cols, h->partial);
  ZF_CUDA(cudaGetLastError());
  zf::zf_count_launch();
  return ZF_OK;
}

// x_new = prox(...) and its sums -> h->sums ; also fetches partial[n_cols] -> *ss
int run_trial(zf_lasso* h, bool abs_only, const double* vec_for_abs, double* ss) {
  if (abs_only)
    zf::lasso_prox_kernel<<<h->vec_blocks, zf::VEC_THREADS, 0, h->st>>>(
        nullptr, nullptr, 0.0, 0.0, 0.0, h->n_cols, const_cast<double*>(vec_for_abs), nullptr,
        h->block_sums, h->counter, h->d_sums);
  else
    zf::lasso_prox_kernel<<<h->vec_blocks, zf::VEC_THREADS, 0, h->st>>>(
        h->y, h->partial, 2.0 * h->scale, h->lr, h->l1 * h->lr, h->n_cols, h->xn, h->g,
        h->block_sums, h->counter, h->d_sums);
  ZF_CUDA(cudaGetLastError());
  zf::zf_count_launch();
  ZF_CUDA(cudaMemcpyAsync(h->h_pin, h->d_sums, sizeof(zf::StepSums), cudaMemcpyDeviceToHost,
                          h->st));
  ZF_CUDA(cudaMemcpyAsync(h->h_pin + 4, h->partial + h->n_cols, sizeof(double),
                          cudaMemcpyDeviceToHost, h->st));
  ZF_CUDA(cudaStreamSynchronize(h->st));
  std::memcpy(&h->sums, h->h_pin, sizeof(zf::StepSums));
  if (ss) *ss = h->h_pin[4];
  return ZF_OK;
}

int fetch_ss(zf_lasso* h, double* ss) {
  ZF_CUDA(cudaMemcpyAsync(h->h_pin + 4, h->partial + h->n_cols, sizeof(double),
                          cudaMemcpyDeviceToHost, h->st));
  ZF_CUDA(cudaStreamSynchronize(h->st));
  *ss = h->h_pin[4];
  return ZF_OK;
}

// f = np.linalg.norm(A @ x - b) ** 2 * scale   (test_proximal_gradient.py:50)
inline double f_from_ss(const zf_lasso* h, double ss) {
  const double nrm = std::sqrt(ss);
  return nrm * nrm * h->scale;
}

// proximal_gradient.py:149-155 for one objective
inline double subproblem_fun(const zf_lasso* h) {
  const double nrm = std::sqrt(h->sums.dd);
  double fun = h->sums.gd + h->l1 * h->sums.abs1 + nrm * nrm / 2.0 / h->lr;
  if (!h->opt.deprecated) fun += h->f_y - h->F_prev;
  return fun;
}

void finish_state(zf_lasso* h, int status) {
  h->status = status;
  h->phase = LP_DONE;
}

// The candidate x_new was accepted by the line search: stop test, momentum, next iterate
// (proximal_gradient.py:510-538).  Returns the next `which` (0 grad, 1 f-eval, 2 done).
int accept_candidate(zf_lasso* h, int* next) {
  h->err = h->sums.maxd;
  const int cap = h->opt.trace_capacity;
  if (cap > 0 && h->nit <= cap) {
    if (h->h_allerrs) h->h_allerrs[h->nit - 1] = h->err;
    if (h->h_allfuns && h->F_known) h->h_allfuns[h->nit] = h->F_x;
  }
  const bool converged = h->err < h->opt.tol;
  if (converged || h->nit >= h->opt.max_iter) {
    h->status = converged ? 1 : 0;
    h->result_is_prev = false;
    if (h->F_known) {
      h->phase = LP_DONE;
      *next = 2;
    } else {
      h->phase = LP_FINAL;   // one more pass for res.fun = F(x)
      *next = 1;
    }
    return ZF_OK;
  }
  double mom = 0.0;
  if (h->opt.nesterov) {
    const double t = h->t_prev;
    const double t_new = std::sqrt(t * t - h->opt.nesterov_a * t + h->opt.nesterov_b) + 0.5;
    mom = (t - 1.0) / t_new;
    h->t_prev = t_new;
  }
  zf::lasso_momentum_kernel<<<h->vec_blocks, zf::VEC_THREADS, 0, h->st>>>(h->xn, h->xp, mom,
                                                                        h->n_cols, h->y);
  ZF_CUDA(cudaGetLastError());
  zf::zf_count_launch();
  double* t = h->xp;
  h->xp = h->xn;
  h->xn = t;
  h->F_prev = h->F_x;
  h->nit += 1;
  h->phase = LP_GRAD;
  *next = 0;
  return ZF_OK;
}

}  // namespace

namespace {

// Configure the chunk-ring gradient pass for clusters of `c` CTAs over rows [0, rows) (the
// handle's full row count, or a prefix while zf_lasso_create times the candidates).  Returns
// false if a row slice does not fit (more than 5 chunks of 2048 columns per CTA).
// 4-CTA clusters leave SMs idle (33 clusters = 132 of 148 SMs on B200).  When a 2-CTA cluster can
// still hold the row slice (<= 5 chunks) and rate2 > 0, a second launch of 2-CTA clusters takes
// the last rows on those SMs, concurrently, on its own stream; its share of the rows follows the
// per-SM rates 0.82 (4-CTA, <= 3 chunks per row) : rate2 (2-CTA, 5 chunks).
// (8-CTA clusters + a 4-CTA second launch was measured: the 4-CTA clusters do not fit on the
// idle SMs and run afterwards, 0.38 instead of 0.59 at 36000 columns -- only c == 4 splits.)
bool ring_plan(zf_lasso* h, int c, double rate2, long long rows, int max_smem) {
  const long long n2 = h->n_cols / 2;
  h->fused_ring = false;
  h->fused_pairs = 0;
  h->ring2_ctas = 0;
  h->ring_rows = rows;
  if (c < 1 || c > zf::RING_MAX_CLUSTER || !h->vec || n2 < c || rows < 1) return false;
  const long long ppc = (n2 + c - 1) / c;
  const long long nchunks = (ppc + zf::RING_CH_PAIRS - 1) / zf::RING_CH_PAIRS;
  if (nchunks > 5 || (size_t)zf::RING_SLOTS * zf::RING_CH_PAIRS * 16 + 4096 > (size_t)max_smem)
    return false;
  h->fused_ring = true;
  h->fused_cluster = c;
  h->fused_pairs_per_cta = ppc;
  h->fused_pairs = (int)nchunks;
  int n_clusters = h->n_sm / c;
  if (n_clusters < 1) { h->fused_ring = false; h->fused_pairs = 0; return false; }
  h->fused_ctas = n_clusters * c;
  h->fused_rows_per_cta = 1;
  int active = 0;
  if (launch_fused_ring(h, nullptr, true, &active) == ZF_OK && active > 0 && active < n_clusters)
    n_clusters = active;
  const int idle = h->n_sm - n_clusters * c;
  // (Only for 4-CTA clusters.  5- and 6-CTA clusters also leave 16-18 SMs idle, but a concurrent
  // 2-CTA launch next to them was measured at 0.40-0.49 ms against 0.24 ms without it on the probe
  // rows of 100000 x 20000: its clusters take SMs the wide clusters need to be co-resident.)
  const int c2 = c / 2;                                  // cluster size of the second launch
  const long long ppc2 = c2 > 0 ? (n2 + c2 - 1) / c2 : n2;
  const long long nch2 = (ppc2 + zf::RING_CH_PAIRS - 1) / zf::RING_CH_PAIRS;
  if (c == 4 && idle >= c2 && nch2 <= 5 && rate2 > 0.0) {
    const int clusters2 = idle / c2;
    const double w1 = 0.82 * n_clusters * c, w2 = rate2 * clusters2 * c2;
    const long long rows2 = (long long)((double)rows * w2 / (w1 + w2));
    bool have = h->st2 != nullptr;
    if (!have)
      have = cudaStreamCreateWithFlags(&h->st2, cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) == cudaSuccess;
    if (rows2 >= clusters2 && have) {
      h->ring2_row0 = rows - rows2;
      h->ring2_rows_per_cluster = (rows2 + clusters2 - 1) / clusters2;
      h->ring2_cluster = c2;
      h->ring2_ctas = c2 * (int)((rows2 + h->ring2_rows_per_cluster - 1) / h->ring2_rows_per_cluster);
      h->ring2_pairs_per_cta = ppc2;
      h->ring2_nch = (int)nch2;
    }
  }
  // rows of the first launch over its clusters
  const long long rows1 = h->ring2_ctas > 0 ? h->ring2_row0 : rows;
  h->fused_rows_per_cta = (rows1 + n_clusters - 1) / n_clusters;
  const int used = (int)((rows1 + h->fused_rows_per_cta - 1) / h->fused_rows_per_cta);
  h->fused_ctas = used * c;
  h->ring1_clusters = used;
  return true;
}

}  // namespace

namespace {

// One-off timed probe at zf_lasso_create (n_cols > 16384, where no single rule fits: how well a
// cluster size does depends on how full its last 2048-column chunk is, on how the clusters pack
// into the GPCs, and -- for the 4-CTA + concurrent 2-CTA form -- on the row share of the second
// launch).  Every candidate plan is timed on the first 64 rows per SM of the caller's own A (two
// launches after a warm-up one, CUDA events) and the fastest is kept; the static choice stays
// unless something beats it by more than 3 %, so that timing noise does not flip the kernel form
// (and with it the summation order of A^T r) between two handles of the same shape.
// ZF_LASSO_TUNE=0 switches the probe off, ZF_LASSO_TUNE=v prints the table.
void ring_autotune(zf_lasso* h, int max_smem) {
  // rps: rows per exchange step (it changes the schedule, not the arithmetic: results are
  // bit-identical to rps 1)
  struct Cand { int c; double rate2; int rps; float ms; int ctas; };
  const int c_static = h->fused_cluster;
  const double r_static = h->ring2_ctas > 0 ? 0.57 : 0.0;
  const int rps_static = h->ring_rps;
  const bool rps_fixed = getenv("ZF_LASSO_RING_RPS") != nullptr;
  std::vector<Cand> cands;
  cands.push_back(Cand{c_static, r_static, rps_static, 0.f, 0});
  const long long n2 = h->n_cols / 2;
  for (int c = 1; c <= zf::RING_MAX_CLUSTER; ++c) {
    const long long ppc = (n2 + c - 1) / c;
    const long long nch = (ppc + zf::RING_CH_PAIRS - 1) / zf::RING_CH_PAIRS;
    if (nch > 4) continue;
    for (double r : {0.0, 0.45, 0.57, 0.7}) {
      if (r > 0.0 && c != 4) continue;
      for (int rps = 1; rps <= 2; ++rps) {
        if (rps_fixed ? rps != rps_static : rps * nch > 6) continue;
        if (c == c_static && r == r_static && (rps == rps_static || nch > 3)) continue;
        cands.push_back(Cand{c, r, rps, 0.f, 0});
      }
    }
  }
  const long long probe_rows = h->n_rows < 64LL * h->n_sm ? h->n_rows : 64LL * h->n_sm;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) {
    if (e0) cudaEventDestroy(e0);
    ring_plan(h, c_static, r_static, h->n_rows, max_smem);
    return;
  }
  cudaMemsetAsync(h->y, 0, sizeof(double) * (size_t)h->n_cols, h->st);
  int best = -1;
  for (size_t k = 0; k < cands.size(); ++k) {
    Cand& cd = cands[k];
    cd.ms = -1.f;
    if (!ring_plan(h, cd.c, cd.rate2, probe_rows, max_smem)) continue;
    if (cd.rate2 > 0.0 && h->ring2_ctas == 0) continue;           // the split did not apply
    h->ring_rps = cd.rps;
    cd.ctas = h->fused_ctas + h->ring2_ctas;
    bool ok = launch_fused_ring(h, h->y, false, nullptr) == ZF_OK;
    ok = ok && cudaEventRecord(e0, h->st) == cudaSuccess;
    for (int rep = 0; rep < 2 && ok; ++rep) ok = launch_fused_ring(h, h->y, false, nullptr) == ZF_OK;
    ok = ok && cudaEventRecord(e1, h->st) == cudaSuccess && cudaEventSynchronize(e1) == cudaSuccess;
    if (!ok) { cudaGetLastError(); continue; }
    cudaEventElapsedTime(&cd.ms, e0, e1);
    cd.ms *= 0.5f;
    if (best < 0 || cd.ms < cands[best].ms) best = (int)k;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  int pick = 0;
  if (best > 0 && (cands[0].ms <= 0.f || cands[best].ms < cands[0].ms / 1.03f)) pick = best;
  if (const char* ev = getenv("ZF_LASSO_TUNE")) {
    if (ev[0] == 'v') {
      for (size_t k = 0; k < cands.size(); ++k)
        fprintf(stderr, "[zf_lasso tune %lldx%lld] cluster %d split %.2f rows/step %d on %3d SMs : %.4f ms%s%s\n",
                h->n_rows, h->n_cols, cands[k].c, cands[k].rate2, cands[k].rps, cands[k].ctas, cands[k].ms,
                k == 0 ? " (static)" : "",
                (int)k == pick ? " <- chosen" : "");
    }
  }
  h->ring_rps = cands[pick].rps;
  if (!ring_plan(h, cands[pick].c, cands[pick].rate2, h->n_rows, max_smem)) {
    h->ring_rps = rps_static;
    ring_plan(h, c_static, r_static, h->n_rows, max_smem);
  }
  h->tuned_cluster = h->fused_cluster;
  h->tuned_rate2 = h->ring2_ctas > 0 ? cands[pick].rate2 : 0.0;
}

}  // namespace

extern "C" int zf_lasso_create(zf_lasso** out, const double* d_A, const double* d_b,
                               int64_t n_rows, int64_t n_cols, double scale, double l1,
                               void* cuda_stream) {
  if (!out || !d_A || !d_b) return zf::zf_fail(ZF_ERR_INVALID, "NULL argument");
  if (n_rows < 1 || n_cols < 1) return zf::zf_fail(ZF_ERR_INVALID, "n_rows and n_cols must be >= 1");
  int rc = zf::zf_require_device();
  if (rc != ZF_OK) return rc;
  zf_lasso* h = new zf_lasso();
  h->A = d_A;
  h->b = d_b;
  h->n_rows = n_rows;
  h->n_cols = n_cols;
  h->scale = scale;
  h->l1 = l1;
  h->st = (cudaStream_t)cuda_stream;
  h->vec = (n_cols % 2 == 0) && ((reinterpret_cast<uintptr_t>(d_A) & 15u) == 0);
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, dev);
  if (h->n_sm < 1) h->n_sm = 148;
  if (const char* e_sm = getenv("ZF_LASSO_NSM")) {     // plan for fewer SMs (tests: a non-148 part)
    const int v_sm = atoi(e_sm);
    if (v_sm >= 1 && v_sm < h->n_sm) h->n_sm = v_sm;
  }
  // residual: 3 CTAs per SM, never more warps than row groups
  const long long n_groups = (n_rows + zf::RES_ROWS_PER_WARP - 1) / zf::RES_ROWS_PER_WARP;
  long long rb = (n_groups + 7) / 8;
  if (rb > 3LL * h->n_sm) rb = 3LL * h->n_sm;
  h->res_blocks = (int)(rb < 1 ? 1 : rb);
  // A^T r: slabs x row blocks ~ 6 CTAs per SM, at least 64 rows per block
  h->n_slabs = (int)((n_cols + zf::ATR_SLAB - 1) / zf::ATR_SLAB);
  long long want = (6LL * h->n_sm + h->n_slabs - 1) / h->n_slabs;
  long long max_rb = (n_rows + 63) / 64;
  if (want > max_rb) want = max_rb;
  if (want < 1) want = 1;
  if (want > 65535) want = 65535;
  h->rows_per_block = (n_rows + want - 1) / want;
  h->n_rowblocks = (int)((n_rows + h->rows_per_block - 1) / h->rows_per_block);
  long long vb = (n_cols + zf::VEC_THREADS - 1) / zf::VEC_THREADS;
  if (vb > zf::VEC_MAX_BLOCKS) vb = zf::VEC_MAX_BLOCKS;
  h->vec_blocks = (int)vb;
  // ---- which kernels compute A^T(A v - b) for this shape.  Measured on B200 (DESIGN.md 3.3,
  // fraction of the one-pass HBM bound).  Warp-specialised chunk ring: 0.89 at 4096 columns,
  // 0.97-0.99 at 6000-8192 (no cluster), 0.95-0.98 at 12000-16384 (cluster 2), 0.82 at 30000
  // (cluster 4: 132 of 148 SMs).  Below 4096 columns the single-CTA fused kernel (0.83 at 4096
  // columns, 0.52 at 6000); two-pass kernels 0.53 everywhere.  (The round-1 cluster / TMA forms,
  // 0.65-0.77, were removed in round 2: nothing selected them any more.)
  // Environment overrides for experiments:
  //   ZF_LASSO_FUSED=0 (two-pass) | 1 (single-CTA fused), ZF_LASSO_RING=1..8 (cluster size),
  //   ZF_LASSO_RING_SPLIT=0 (no second launch on the idle SMs) | .NN (its per-SM rate),
  //   ZF_LASSO_TUNE=0|v, ZF_LASSO_NSM=n.
  h->gpart_rows = (size_t)h->n_rowblocks;
  size_t sq_rows = (size_t)h->res_blocks;
  {
    const long long n2 = n_cols / 2;
    const long long pairs = (n2 + zf::FUSED_THREADS - 1) / zf::FUSED_THREADS;  // per thread, 1 CTA
    int max_smem = 0;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const bool big_enough = h->vec && n_rows >= 4LL * h->n_sm;
    const char* env_fused = getenv("ZF_LASSO_FUSED");
    const bool off = env_fused && env_fused[0] == '0';
    if (const char* env_rps = getenv("ZF_LASSO_RING_RPS")) h->ring_rps = atoi(env_rps) == 2 ? 2 : 1;
    auto try_ring = [&](int c) -> bool {
      if (!big_enough) return false;
      const char* env_split = getenv("ZF_LASSO_RING_SPLIT");
      double rate2 = 0.57;       // 200000 x 20000: none 0.851, .5: 0.913, .55: 0.920, .6: 0.926,
                                 // .65: 0.872, .7: 0.815 (second launch becomes the long pole)
      if (env_split && env_split[0] == '0') rate2 = 0.0;
      else if (env_split && atof(env_split) > 0.0) rate2 = atof(env_split);   // experiment knob
      if (!ring_plan(h, c, rate2, n_rows, max_smem)) return false;
      const size_t parts = (size_t)h->ring1_clusters + (size_t)(h->ring2_ctas / h->ring2_cluster);
      if (parts > h->gpart_rows) h->gpart_rows = parts;
      if (parts > sq_rows) sq_rows = parts;
      return true;
    };
    auto try_single = [&]() -> bool {
      if (!big_enough || pairs > 20 || (size_t)n_cols * 8 + 1024 > (size_t)max_smem) return false;
      h->fused_pairs = (int)(((pairs + 3) / 4) * 4);
      h->fused_ctas = h->n_sm;
      h->fused_rows_per_cta = (n_rows + h->fused_ctas - 1) / h->fused_ctas;
      if (h->fused_rows_per_cta % 2) h->fused_rows_per_cta += 1;     // whole row pairs
      h->fused_ctas = (int)((n_rows + h->fused_rows_per_cta - 1) / h->fused_rows_per_cta);
      if ((size_t)h->fused_ctas > h->gpart_rows) h->gpart_rows = (size_t)h->fused_ctas;
      if ((size_t)h->fused_ctas > sq_rows) sq_rows = (size_t)h->fused_ctas;
      return true;
    };
    const char* env_ring = getenv("ZF_LASSO_RING");
    if (off) {
      // two-pass kernels
    } else if (env_ring) {
      try_ring(atoi(env_ring));
    } else if (env_fused) {
      try_single();
    } else if (n_cols < 4096) {
      try_single();
    } else {
      // chunk ring: the smallest cluster whose row slice is at most 4 chunks (8192 columns); 7-CTA
      // clusters pack badly into the GPCs and are never the static choice (sustained, B200:
      // 36000 columns 5 / 6 / 8 CTAs 0.89 / 0.92 / 0.70, 40000: 0.97 / 0.81 / 0.77, 50000 and
      // 60000: 8 CTAs 0.73 / 0.85).  Above 16384 columns the create-time probe below re-decides.
      const int c = n_cols <= 8192 ? 1 : n_cols <= 16384 ? 2 : n_cols <= 32768 ? 4
                  : n_cols <= 40960 ? 5 : n_cols <= 49152 ? 6 : 8;
      // (rows wider than 8 x 5 chunks = 81920 columns, an odd column count or an unaligned A:
      // the two-pass kernels)
      if (!try_ring(c) && !try_ring(8)) try_single();
    }
  }
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void** p, size_t bytes) {
    if (e == cudaSuccess) e = cudaMalloc(p, bytes);
  };
  const bool tune = h->fused_ring && n_cols > 16384 && !getenv("ZF_LASSO_RING") &&
                    !(getenv("ZF_LASSO_TUNE") && getenv("ZF_LASSO_TUNE")[0] == '0');
  if (tune) {                     // any candidate plan leaves at most one partial row per SM
    if ((size_t)h->n_sm > h->gpart_rows) h->gpart_rows = (size_t)h->n_sm;
    if ((size_t)h->n_sm > sq_rows) sq_rows = (size_t)h->n_sm;
  }
  alloc((void**)&h->vecs, sizeof(double) * 4 * (size_t)n_cols);
  alloc((void**)&h->r, sizeof(double) * (size_t)n_rows);
  alloc((void**)&h->gpart, sizeof(double) * h->gpart_rows * (size_t)n_cols);
  alloc((void**)&h->sq_part, sizeof(double) * sq_rows);
  alloc((void**)&h->partial, sizeof(double) * ((size_t)n_cols + 1));
  alloc((void**)&h->block_sums, sizeof(zf::StepSums) * zf::VEC_MAX_BLOCKS);
  alloc((void**)&h->d_sums, sizeof(zf::StepSums));
  alloc((void**)&h->counter, sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMallocHost((void**)&h->h_pin, sizeof(double) * 8);
  alloc((void**)&h->d_opts, sizeof(zf::LassoDevOpts));
  alloc((void**)&h->d_state, sizeof(zf::LassoDevState));
  if (e == cudaSuccess) e = cudaMemsetAsync(h->d_state, 0, sizeof(zf::LassoDevState), h->st);
  if (e == cudaSuccess) e = cudaMallocHost((void**)&h->h_state, 2 * sizeof(zf::LassoDevState));
  for (int k = 0; k < 2 && e == cudaSuccess; ++k)
    e = cudaEventCreateWithFlags(&h->ev_poll[k], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaMemsetAsync(h->counter, 0, sizeof(unsigned int), h->st);
  if (e != cudaSuccess) {
    zf_lasso_destroy(h);
    return zf::zf_fail_cuda(e, "zf_lasso_create allocation");
  }
  h->xp = h->vecs;
  h->xn = h->vecs + n_cols;
  h->y = h->vecs + 2 * n_cols;
  h->g = h->vecs + 3 * n_cols;
  if (tune) {
    int max_smem = 0;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    ring_autotune(h, max_smem);
  }
  *out = h;
  return ZF_OK;
}

extern "C" void zf_lasso_destroy(zf_lasso* h) {
  if (!h) return;
  cudaFree(h->vecs);
  cudaFree(h->r);
  cudaFree(h->gpart);
  cudaFree(h->sq_part);
  cudaFree(h->partial);
  cudaFree(h->block_sums);
  cudaFree(h->d_sums);
  cudaFree(h->counter);
  if (h->h_pin) cudaFreeHost(h->h_pin);
  for (int r = 0; r < zf::P2P_MAX_RANKS; ++r)
    if (h->p2p_opened[r]) cudaIpcCloseMemHandle(h->p2p_opened[r]);
  cudaFree(h->p2p_mine);
  cudaFree(h->ss_gather);
  cudaFree(h->p2p_ticket);
  cudaFree(h->d_opts);
  cudaFree(h->d_state);
  cudaFree(h->d_allerrs);
  cudaFree(h->d_allfuns);
  cudaFree(h->d_allvecs);
  if (h->h_state) cudaFreeHost(h->h_state);
  for (int k = 0; k < 2; ++k) {
    if (h->ev_poll[k]) cudaEventDestroy(h->ev_poll[k]);
    if (h->graph[k]) cudaGraphExecDestroy(h->graph[k]);
  }
  if (h->ev_own) cudaEventDestroy(h->ev_own);
  if (h->st_own) cudaStreamDestroy(h->st_own);
  if (h->st2) cudaStreamDestroy(h->st2);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  delete h;
}

static int lasso_check_options(const zf_options* o) {
  if (!o) return zf::zf_fail(ZF_ERR_INVALID, "options is NULL");
  if (!(o->lr > 0.0)) return zf::zf_fail(ZF_ERR_INVALID, "lr must be > 0");
  if (o->max_iter < 1) return zf::zf_fail(ZF_ERR_INVALID, "max_iter must be >= 1");
  if (o->max_backtrack_iter < 1) return zf::zf_fail(ZF_ERR_INVALID, "max_backtrack_iter must be >= 1");
  if (!(o->decay_rate > 0.0 && o->decay_rate <= 1.0))
    return zf::zf_fail(ZF_ERR_INVALID, "decay_rate must be in (0, 1]");
  if (o->trace_capacity < 0) return zf::zf_fail(ZF_ERR_INVALID, "trace_capacity must be >= 0");
  return ZF_OK;
}

static int lasso_begin_impl(zf_lasso* h, const zf_options* opt, const double* d_x0,
                            double* h_allerrs, double* h_allfuns) {
  if (!h || !d_x0) return zf::zf_fail(ZF_ERR_INVALID, "NULL argument");
  int rc = lasso_check_options(opt);
  if (rc != ZF_OK) return rc;
  h->opt = *opt;
  h->h_allerrs = h_allerrs;
  h->h_allfuns = h_allfuns;
  h->lr = opt->lr;
  h->t_prev = 1.0;
  h->nit = 0;
  h->status = 0;
  h->bt = 0;
  h->err = INFINITY;
  h->result_is_prev = false;
  // F is needed by the line search, and by return_all's allfuns
  h->need_F = (opt->decay_rate != 1.0) || (opt->trace_capacity > 0 && h_allfuns != nullptr);
  h->F_known = false;
  h->xp = h->vecs;
  h->xn = h->vecs + h->n_cols;
  const size_t nb = sizeof(double) * (size_t)h->n_cols;
  ZF_CUDA(cudaMemcpyAsync(h->xp, d_x0, nb, cudaMemcpyDeviceToDevice, h->st));
  ZF_CUDA(cudaMemcpyAsync(h->xn, d_x0, nb, cudaMemcpyDeviceToDevice, h->st));
  ZF_CUDA(cudaMemcpyAsync(h->y, d_x0, nb, cudaMemcpyDeviceToDevice, h->st));
  // F(x0): residual norm partial into `partial[n_cols]` (all-reduced by the caller if sharded)
  rc = launch_residual(h, h->xp);
  if (rc != ZF_OK) return rc;
  rc = launch_collect(h, false);
  if (rc != ZF_OK) return rc;
  h->phase = LP_INIT;
  return ZF_OK;
}

extern "C" int zf_lasso_begin(zf_lasso* h, const zf_options* opt, const double* d_x0) {
  return lasso_begin_impl(h, opt, d_x0, nullptr, nullptr);
}

extern "C" int zf_lasso_grad(zf_lasso* h, int which) {
  if (!h) return zf::zf_fail(ZF_ERR_INVALID, "NULL handle");
  int rc;
  if (which == 0) {
    if (h->phase != LP_GRAD) return zf::zf_fail(ZF_ERR_INVALID, "zf_lasso_grad(0) out of order");
    int ng = 0, ns = 0;
    rc = launch_gradient_pass(h, h->y, &ng, &ns);
    if (rc != ZF_OK) return rc;
    return launch_collect_n(h, true, ng, ns);
  }
  if (which == 1) {
    if (h->phase != LP_FNEW && h->phase != LP_FINAL)
      return zf::zf_fail(ZF_ERR_INVALID, "zf_lasso_grad(1) out of order");
    rc = launch_residual(h, h->xn);
    if (rc != ZF_OK) return rc;
    return launch_collect(h, false);
  }
  return zf::zf_fail(ZF_ERR_INVALID, "which must be 0 or 1");
}

extern "C" double* zf_lasso_partial(zf_lasso* h, int64_t* n_values) {
  if (!h) return nullptr;
  if (n_values) *n_values = h->n_cols + 1;
  return h->partial;
}

extern "C" int zf_lasso_step(zf_lasso* h, int32_t* h_next) {
  if (!h || !h_next) return zf::zf_fail(ZF_ERR_INVALID, "NULL argument");
  int next = 2;
  int rc = ZF_OK;
  double ss = 0.0;
  switch (h->phase) {
    case LP_INIT: {
      // F(x0) = f(x0) + g(x0)   (proximal_gradient.py:466, 279)
      rc = run_trial(h, true, h->xp, &ss);
      if (rc != ZF_OK) return rc;
      h->F_prev = f_from_ss(h, ss) + h->l1 * h->sums.abs1;
      h->F_x = h->F_prev;
      if (h->opt.trace_capacity > 0 && h->h_allfuns) h->h_allfuns[0] = h->F_prev;
      h->nit = 1;
      h->phase = LP_GRAD;
      next = 0;
      break;
    }
    case LP_GRAD: {
      h->bt = 0;
      rc = run_trial(h, false, nullptr, &ss);
      if (rc != ZF_OK) return rc;
      h->f_y = f_from_ss(h, ss);
      h->sub_fun = subproblem_fun(h);
      if (h->need_F) {
        h->phase = LP_FNEW;
        next = 1;
      } else {
        h->F_known = false;
        rc = accept_candidate(h, &next);
      }
      break;
    }
    case LP_FNEW: {
      rc = fetch_ss(h, &ss);
      if (rc != ZF_OK) return rc;
      const double f_x = f_from_ss(h, ss);
      h->F_x = f_x + h->l1 * h->sums.abs1;
      h->F_known = true;
      bool ok;
      if (h->opt.decay_rate == 1.0) ok = true;                      // proximal_gradient.py:298
      else if (h->opt.deprecated) ok = (f_x - h->f_y <= h->sub_fun + h->opt.tol_internal);
      else ok = (h->F_x - h->F_prev <= h->sub_fun + h->opt.tol_internal);
      if (ok) {
        rc = accept_candidate(h, &next);
      } else {
        h->lr *= h->opt.decay_rate;
        h->bt += 1;
        if (h->bt >= h->opt.max_backtrack_iter) {
          // RuntimeError("Backtracking failed ...") -> x = x_prev, nit - 1 (493-509)
          h->result_is_prev = true;
          h->F_x = h->F_prev;
          h->nit -= 1;
          finish_state(h, -1);
          next = 2;
        } else {
          // same gradient, smaller step: redo the prox from the stored g
          zf::lasso_prox_kernel<<<h->vec_blocks, zf::VEC_THREADS, 0, h->st>>>(
              h->y, h->g, 1.0, h->lr, h->l1 * h->lr, h->n_cols, h->xn, nullptr, h->block_sums,
              h->counter, h->d_sums);
          ZF_CUDA(cudaGetLastError());
          zf::zf_count_launch();
          ZF_CUDA(cudaMemcpyAsync(h->h_pin, h->d_sums, sizeof(zf::StepSums),
                                  cudaMemcpyDeviceToHost, h->st));
          ZF_CUDA(cudaStreamSynchronize(h->st));
          std::memcpy(&h->sums, h->h_pin, sizeof(zf::StepSums));
          h->sub_fun = subproblem_fun(h);
          h->phase = LP_FNEW;
          next = 1;
        }
      }
      break;
    }
    case LP_FINAL: {
      rc = fetch_ss(h, &ss);
      if (rc != ZF_OK) return rc;
      h->F_x = f_from_ss(h, ss) + h->l1 * h->sums.abs1;
      h->F_known = true;
      h->phase = LP_DONE;
      next = 2;
      break;
    }
    case LP_DONE:
      next = 2;
      break;
    default:
      return zf::zf_fail(ZF_ERR_INVALID, "zf_lasso_step called before zf_lasso_begin");
  }
  if (rc != ZF_OK) return rc;
  *h_next = next;
  return ZF_OK;
}

extern "C" int zf_lasso_finish(zf_lasso* h, double* d_x, double* h_fun, int64_t* h_nit,
                               int32_t* h_status) {
  if (!h) return zf::zf_fail(ZF_ERR_INVALID, "NULL handle");
  if (h->phase != LP_DONE) return zf::zf_fail(ZF_ERR_INVALID, "zf_lasso_finish before the solve ended");
  if (d_x) {
    ZF_CUDA(cudaMemcpyAsync(d_x, h->result_is_prev ? h->xp : h->xn,
                            sizeof(double) * (size_t)h->n_cols, cudaMemcpyDeviceToDevice, h->st));
    ZF_CUDA(cudaStreamSynchronize(h->st));
  }
  if (h_fun) *h_fun = h->F_x;
  if (h_nit) *h_nit = h->nit;
  if (h_status) *h_status = h->status;
  h->phase = LP_IDLE;
  return ZF_OK;
}

static int lasso_solve_dev(zf_lasso* h, const zf_options* opt, const double* d_x0, double* d_x,
                           double* h_fun, int64_t* h_nit, int32_t* h_status, double* h_allerrs,
                           double* h_allfuns);

extern "C" int zf_lasso_solve(zf_lasso* h, const zf_options* opt, const double* d_x0,
                              double* d_x, double* h_fun, int64_t* h_nit, int32_t* h_status,
                              double* h_allerrs, double* h_allfuns) {
  // default: the device-decided loop; ZF_LASSO_HOSTLOOP=1 keeps the round-1 loop (one D2H copy +
  // stream sync per trial, decisions on the host) for A/B measurements
  static const bool hostloop = getenv("ZF_LASSO_HOSTLOOP") && getenv("ZF_LASSO_HOSTLOOP")[0] == '1';
  if (!hostloop) {
    if (!h) return zf::zf_fail(ZF_ERR_INVALID, "NULL handle");
    return lasso_solve_dev(h, opt, d_x0, d_x, h_fun, h_nit, h_status, h_allerrs, h_allfuns);
  }
  h->h_allvecs = nullptr;          // the host-decided loop does not record iterates
  h->h_allvecs_rows = 0;
  int rc = lasso_begin_impl(h, opt, d_x0, h_allerrs, h_allfuns);
  if (rc != ZF_OK) return rc;
  int32_t next = 0;
  rc = zf_lasso_step(h, &next);
  while (rc == ZF_OK && next != 2) {
    rc = zf_lasso_grad(h, next);
    if (rc != ZF_OK) break;
    rc = zf_lasso_step(h, &next);
  }
  if (rc != ZF_OK) {
    h->phase = LP_IDLE;
    return rc;
  }
  return zf_lasso_finish(h, d_x, h_fun, h_nit, h_status);
}

extern "C" int zf_lasso_passes(zf_lasso* h) {
  if (!h) return 0;
  return h->fused_pairs > 0 ? 1 : 2;
}

extern "C" int zf_lasso_gradient_device(zf_lasso* h, const double* d_x, double* d_grad,
                                        double* d_f) {
  if (!h || !d_x || !d_grad) return zf::zf_fail(ZF_ERR_INVALID, "NULL argument");
  int ng = 0, ns = 0;
  int rc = launch_gradient_pass(h, d_x, &ng, &ns);
  if (rc != ZF_OK) return rc;
  rc = launch_collect_n(h, true, ng, ns);
  if (rc != ZF_OK) return rc;
  zf::lasso_scale_kernel<<<h->vec_blocks, zf::VEC_THREADS, 0, h->st>>>(
      h->partial, 2.0 * h->scale, h->scale, h->n_cols, d_grad, d_f);
  ZF_CUDA(cudaGetLastError());
  zf::zf_count_launch();
  return ZF_OK;
}

// =======================================================================================
// device-decided loop: host side (enqueue only; see the kernels above)
// =======================================================================================
namespace {

enum { DS_INIT = 0, DS_GRAD = 1, DS_PROX = 2, DS_FEVAL = 3, DS_DECIDE = 4, DS_FINAL = 5,
       DS_FEVAL_FINAL = 6 };

// how many partial rows / residual-norm partials the handle's gradient pass leaves behind
void gradient_part_counts(const zf_lasso* h, int* n_parts, int* n_sq) {
  if (h->fused_pairs > 0 && h->fused_ring) {
    *n_parts = h->fused_ctas / h->fused_cluster + h->ring2_ctas / h->ring2_cluster;
    *n_sq = *n_parts;
  } else if (h->fused_pairs > 0) {
    *n_parts = h->fused_ctas;
    *n_sq = h->fused_ctas;
  } else {
    *n_parts = h->n_rowblocks;
    *n_sq = h->res_blocks;
  }
}

template <int SRC, bool FIXED>
int launch_dev_update(zf_lasso* h, int n_parts, int n_sq) {
  zf::lasso_dev_update_kernel<SRC, FIXED><<<h->vec_blocks, zf::VEC_THREADS, 0, h->st>>>(
      h->d_opts, h->d_state, h->gpart, n_parts, h->sq_part, n_sq, h->partial, h->n_cols, h->y,
      h->xp, h->xn, h->g, h->block_sums, h->pp);
  ZF_CUDA(cudaGetLastError());
  zf::zf_count_launch();
  return ZF_OK;
}

// publish this rank's sum r^2 and collect every rank's (peer exchange)
int p2p_exchange_ss(zf_lasso* h, int force) {
  zf::lasso_p2p_publish_ss_kernel<<<1, 32, 0, h->st>>>(h->d_state, h->sq_part, h->res_blocks, h->pp, force);
  ZF_CUDA(cudaGetLastError());
  zf::zf_count_launch();
  zf::lasso_p2p_gather_ss_kernel<<<1, 32, 0, h->st>>>(h->d_state, h->pp, h->ss_gather, force);
  ZF_CUDA(cudaGetLastError());
  zf::zf_count_launch();
  return ZF_OK;
}

int dev_stage(zf_lasso* h, int stage) {
  int rc = ZF_OK, ng = 0, ns = 0;
  const bool p2p = h->dev_sharded && h->p2p_on;
  const double* ss_feval = p2p ? h->ss_gather : h->dev_sharded ? h->partial + h->n_cols : h->sq_part;
  const int n_feval = p2p ? h->pp.world : h->dev_sharded ? 1 : h->res_blocks;
  switch (stage) {
    case DS_INIT:
      zf::lasso_dev_init_kernel<<<1, 32, 0, h->st>>>(h->d_opts, h->d_state, h->d_sums, ss_feval,
                                                    n_feval);
      ZF_CUDA(cudaGetLastError());
      zf::zf_count_launch();
      return ZF_OK;
    case DS_GRAD:
      h->skip = &h->d_state->skip_grad;
      rc = launch_gradient_pass(h, h->y, &ng, &ns);
      h->skip = nullptr;
      if (rc != ZF_OK) return rc;
      if (p2p) {                      // collect into this rank's exchange buffer and raise its flag
        const int blocks = (int)((h->n_cols + 255) / 256);
        zf::lasso_p2p_publish_kernel<<<blocks, 256, 0, h->st>>>(h->d_state, h->gpart, ng, h->sq_part,
                                                               ns, h->n_cols, h->pp, h->p2p_ticket);
        ZF_CUDA(cudaGetLastError());
        zf::zf_count_launch();
        return ZF_OK;
      }
      if (h->dev_sharded) rc = launch_collect_n(h, true, ng, ns);   // -> partial, all-reduced next
      return rc;
    case DS_PROX:
      gradient_part_counts(h, &ng, &ns);
      if (p2p)
        return h->dev_fixed ? launch_dev_update<2, true>(h, ng, ns) : launch_dev_update<2, false>(h, ng, ns);
      if (h->dev_sharded)
        return h->dev_fixed ? launch_dev_update<1, true>(h, ng, ns) : launch_dev_update<1, false>(h, ng, ns);
      return h->dev_fixed ? launch_dev_update<0, true>(h, ng, ns) : launch_dev_update<0, false>(h, ng, ns);
    case DS_FEVAL:
    case DS_FEVAL_FINAL:
      h->skip = (stage == DS_FEVAL) ? &h->d_state->done : nullptr;
      rc = launch_residual(h, h->xn);
      h->skip = nullptr;
      if (rc != ZF_OK) return rc;
      if (p2p) return p2p_exchange_ss(h, stage == DS_FEVAL_FINAL ? 1 : 0);
      if (h->dev_sharded) rc = launch_collect(h, false);            // sum r^2 -> partial[n_cols]
      return rc;
    case DS_DECIDE:
      zf::lasso_dev_decide_kernel<<<1, 32, 0, h->st>>>(h->d_opts, h->d_state, ss_feval, n_feval);
      ZF_CUDA(cudaGetLastError());
      zf::zf_count_launch();
      zf::lasso_dev_momentum_kernel<<<h->vec_blocks, zf::VEC_THREADS, 0, h->st>>>(
          h->d_opts, h->d_state, h->n_cols, h->xn, h->xp, h->y);
      ZF_CUDA(cudaGetLastError());
      zf::zf_count_launch();
      return ZF_OK;
    case DS_FINAL:
      zf::lasso_dev_final_kernel<<<1, 32, 0, h->st>>>(h->d_opts, h->d_state, ss_feval, n_feval);
      ZF_CUDA(cudaGetLastError());
      zf::zf_count_launch();
      return ZF_OK;
    default:
      return zf::zf_fail(ZF_ERR_INVALID, "unknown stage");
  }
}

// one trial on one GPU (no exchange between the stages)
int dev_slot(zf_lasso* h) {
  int rc = dev_stage(h, DS_GRAD);
  if (rc == ZF_OK) rc = dev_stage(h, DS_PROX);
  if (rc == ZF_OK && !h->dev_fixed) {
    rc = dev_stage(h, DS_FEVAL);
    if (rc == ZF_OK) rc = dev_stage(h, DS_DECIDE);
  }
  return rc;
}

int dev_snapshot(zf_lasso* h, int slot) {
  ZF_CUDA(cudaMemcpyAsync(&h->h_state[slot], h->d_state, sizeof(zf::LassoDevState),
                          cudaMemcpyDeviceToHost, h->st));
  ZF_CUDA(cudaEventRecord(h->ev_poll[slot], h->st));
  return ZF_OK;
}

constexpr int DEV_GRAPH_SLOTS = 32;

// `DEV_GRAPH_SLOTS` slots as one CUDA graph (captured once per handle and mode; every pointer a
// slot touches is fixed for the life of the handle, the options live in device memory)
int dev_build_graph(zf_lasso* h) {
  const int which = h->dev_fixed ? 0 : 1;
  if (h->graph[which] || h->graph_failed) return ZF_OK;
  static const bool off = getenv("ZF_LASSO_GRAPH") && getenv("ZF_LASSO_GRAPH")[0] == '0';
  if (off) { h->graph_failed = true; return ZF_OK; }
  // warm-up outside the capture: attribute calls and lazy module loading happen here
  cudaGraph_t g = nullptr;
  if (cudaStreamBeginCapture(h->st, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
    cudaGetLastError();
    h->graph_failed = true;
    return ZF_OK;
  }
  int rc = ZF_OK;
  for (int k = 0; k < DEV_GRAPH_SLOTS && rc == ZF_OK; ++k) rc = dev_slot(h);
  const cudaError_t e = cudaStreamEndCapture(h->st, &g);
  if (rc != ZF_OK || e != cudaSuccess || !g ||
      cudaGraphInstantiate(&h->graph[which], g, 0) != cudaSuccess) {
    cudaGetLastError();
    h->graph[which] = nullptr;
    h->graph_failed = true;
  }
  if (g) cudaGraphDestroy(g);
  h->graph_slots = DEV_GRAPH_SLOTS;
  return ZF_OK;
}

}  // namespace

extern "C" int zf_lasso_set_stream(zf_lasso* h, void* cuda_stream) {
  if (!h) return zf::zf_fail(ZF_ERR_INVALID, "NULL handle");
  if (h->dev_active) return zf::zf_fail(ZF_ERR_INVALID, "zf_lasso_set_stream during a solve");
  if ((cudaStream_t)cuda_stream != h->st) {
    for (int k = 0; k < 2; ++k) {        // graphs were captured on the old stream's fork / join
      if (h->graph[k]) cudaGraphExecDestroy(h->graph[k]);
      h->graph[k] = nullptr;
    }
    h->graph_failed = false;
  }
  h->st = (cudaStream_t)cuda_stream;
  return ZF_OK;
}

extern "C" int zf_lasso_dev_begin(zf_lasso* h, const zf_options* opt, const double* d_x0,
                                  int32_t sharded, int32_t want_trace) {
  if (!h || !d_x0) return zf::zf_fail(ZF_ERR_INVALID, "NULL argument");
  int rc = lasso_check_options(opt);
  if (rc != ZF_OK) return rc;
  h->opt = *opt;
  h->dev_sharded = sharded != 0;
  const int cap = want_trace ? opt->trace_capacity : 0;
  if (cap > h->dev_cap) {
    cudaFree(h->d_allerrs);
    cudaFree(h->d_allfuns);
    h->d_allerrs = h->d_allfuns = nullptr;
    h->dev_cap = 0;
    ZF_CUDA(cudaMalloc((void**)&h->d_allerrs, sizeof(double) * (size_t)cap));
    ZF_CUDA(cudaMalloc((void**)&h->d_allfuns, sizeof(double) * ((size_t)cap + 1)));
    h->dev_cap = cap;
  }
  zf::LassoDevOpts& o = h->dev_opts_host;
  o.lr0 = opt->lr;
  o.tol = opt->tol;
  o.tol_internal = opt->tol_internal;
  o.decay = opt->decay_rate;
  o.na = opt->nesterov_a;
  o.nb = opt->nesterov_b;
  o.scale = h->scale;
  o.l1 = h->l1;
  o.max_iter = opt->max_iter;
  o.max_bt = opt->max_backtrack_iter;
  o.nesterov = opt->nesterov;
  o.deprecated = opt->deprecated;
  o.need_F = (opt->decay_rate != 1.0) || cap > 0;     // the line search, or return_all's allfuns
  o.cap = cap;
  o.allerrs = cap > 0 ? h->d_allerrs : nullptr;
  o.allfuns = cap > 0 ? h->d_allfuns : nullptr;
  o.allvecs = nullptr;
  o.vec_rows = 0;
  if (cap > 0 && h->h_allvecs && h->h_allvecs_rows > 0) {
    // iterates x^0 .. x^{rows-1} (return_all's allvecs); rows is bounded by the caller's buffer
    long long rows = h->h_allvecs_rows < (long long)cap + 1 ? h->h_allvecs_rows : (long long)cap + 1;
    if (rows > h->d_allvecs_rows) {
      cudaFree(h->d_allvecs);
      h->d_allvecs = nullptr;
      h->d_allvecs_rows = 0;
      ZF_CUDA(cudaMalloc((void**)&h->d_allvecs, sizeof(double) * (size_t)rows * (size_t)h->n_cols));
      h->d_allvecs_rows = rows;
    }
    o.allvecs = h->d_allvecs;
    o.vec_rows = rows;
    ZF_CUDA(cudaMemcpyAsync(h->d_allvecs, d_x0, sizeof(double) * (size_t)h->n_cols,
                            cudaMemcpyDeviceToDevice, h->st));
  }
  h->dev_fixed = !o.need_F;
  h->xp = h->vecs;
  h->xn = h->vecs + h->n_cols;
  // (pageable source: the runtime stages it before returning, so o may change afterwards)
  ZF_CUDA(cudaMemcpyAsync(h->d_opts, &o, sizeof(o), cudaMemcpyHostToDevice, h->st));
  const size_t nb = sizeof(double) * (size_t)h->n_cols;
  ZF_CUDA(cudaMemcpyAsync(h->xp, d_x0, nb, cudaMemcpyDeviceToDevice, h->st));
  ZF_CUDA(cudaMemcpyAsync(h->xn, d_x0, nb, cudaMemcpyDeviceToDevice, h->st));
  ZF_CUDA(cudaMemcpyAsync(h->y, d_x0, nb, cudaMemcpyDeviceToDevice, h->st));
  // F(x0): residual norm of x0 and ||x0||_1
  rc = launch_residual(h, h->xp);
  if (rc != ZF_OK) return rc;
  zf::lasso_prox_kernel<<<h->vec_blocks, zf::VEC_THREADS, 0, h->st>>>(
      nullptr, nullptr, 0.0, 0.0, 0.0, h->n_cols, h->xp, nullptr, h->block_sums, h->counter,
      h->d_sums);
  ZF_CUDA(cudaGetLastError());
  zf::zf_count_launch();
  if (h->dev_sharded && h->p2p_on) {
    rc = p2p_exchange_ss(h, 1);
    if (rc != ZF_OK) return rc;
  } else if (h->dev_sharded) {
    rc = launch_collect(h, false);
    if (rc != ZF_OK) return rc;
  }
  h->dev_active = true;
  h->phase = LP_IDLE;
  return ZF_OK;
}

/* stage: 0 F(x0) | 1 gradient pass at y | 2 prox / update | 3 residual at the candidate |
 * 4 decide + momentum | 5 res.fun | 6 residual at the final x (before 5 when stage 3/4 never ran) */
extern "C" int zf_lasso_dev_stage(zf_lasso* h, int32_t stage) {
  if (!h || !h->dev_active) return zf::zf_fail(ZF_ERR_INVALID, "zf_lasso_dev_stage before zf_lasso_dev_begin");
  return dev_stage(h, stage);
}

/* 1 when every trial needs stages 3 and 4 (line search or F trace), 0 for the fixed-step loop */
extern "C" int zf_lasso_dev_needs_feval(zf_lasso* h) {
  return (h && h->dev_active && !h->dev_fixed) ? 1 : 0;
}

/* enqueue `n_slots` whole trials (single GPU; as CUDA graph launches where possible) */
extern "C" int zf_lasso_dev_slots(zf_lasso* h, int32_t n_slots) {
  if (!h || !h->dev_active) return zf::zf_fail(ZF_ERR_INVALID, "zf_lasso_dev_slots before zf_lasso_dev_begin");
  if (h->dev_sharded) return zf::zf_fail(ZF_ERR_INVALID, "a row-sharded run enqueues stage by stage");
  int rc = ZF_OK;
  const int which = h->dev_fixed ? 0 : 1;
  if (!h->graph[which] && !h->graph_failed && n_slots > 0) {
    // the first trial of a mode runs outside the capture: its kernels get loaded and their
    // attributes set by ordinary launches
    rc = dev_slot(h);
    if (rc != ZF_OK) return rc;
    n_slots -= 1;
    rc = dev_build_graph(h);
    if (rc != ZF_OK) return rc;
  }
  while (n_slots > 0) {
    if (h->graph[which] && n_slots >= h->graph_slots) {
      ZF_CUDA(cudaGraphLaunch(h->graph[which], h->st));
      zf::zf_count_launch();
      n_slots -= h->graph_slots;
    } else {
      rc = dev_slot(h);
      if (rc != ZF_OK) return rc;
      n_slots -= 1;
    }
  }
  return ZF_OK;
}

/* snapshot the device state after everything enqueued so far (slot 0 / 1), or wait for an
 * earlier snapshot and read it: done != 0 once the solve has ended */
extern "C" int zf_lasso_dev_poll(zf_lasso* h, int32_t slot, int32_t wait, int32_t* h_done,
                                 int64_t* h_nit) {
  if (!h || !h->dev_active || slot < 0 || slot > 1) return zf::zf_fail(ZF_ERR_INVALID, "bad poll");
  if (!wait) return dev_snapshot(h, slot);
  ZF_CUDA(cudaEventSynchronize(h->ev_poll[slot]));
  if (h_done) *h_done = h->h_state[slot].done;
  if (h_nit) *h_nit = h->h_state[slot].nit;
  return ZF_OK;
}

extern "C" int zf_lasso_dev_finish(zf_lasso* h, double* d_x, double* h_fun, int64_t* h_nit,
                                   int32_t* h_status, double* h_lr, double* h_allerrs,
                                   double* h_allfuns) {
  if (!h || !h->dev_active) return zf::zf_fail(ZF_ERR_INVALID, "zf_lasso_dev_finish before zf_lasso_dev_begin");
  h->dev_active = false;
  ZF_CUDA(cudaMemcpyAsync(&h->h_state[0], h->d_state, sizeof(zf::LassoDevState),
                          cudaMemcpyDeviceToHost, h->st));
  ZF_CUDA(cudaStreamSynchronize(h->st));
  const zf::LassoDevState& s = h->h_state[0];
  if (!s.done) return zf::zf_fail(ZF_ERR_INVALID, "zf_lasso_dev_finish before the solve ended");
  if (d_x) {
    ZF_CUDA(cudaMemcpyAsync(d_x, s.result_is_prev ? h->xp : h->xn,
                            sizeof(double) * (size_t)h->n_cols, cudaMemcpyDeviceToDevice, h->st));
  }
  const long long k = s.nit < (long long)h->dev_opts_host.cap ? s.nit : (long long)h->dev_opts_host.cap;
  if (h_allerrs && h->dev_opts_host.cap > 0 && k > 0)
    ZF_CUDA(cudaMemcpyAsync(h_allerrs, h->d_allerrs, sizeof(double) * (size_t)k,
                            cudaMemcpyDeviceToHost, h->st));
  if (h_allfuns && h->dev_opts_host.cap > 0)
    ZF_CUDA(cudaMemcpyAsync(h_allfuns, h->d_allfuns, sizeof(double) * (size_t)(k + 1),
                            cudaMemcpyDeviceToHost, h->st));
  if (h->dev_opts_host.allvecs && h->h_allvecs) {
    long long rows = s.nit + 1 < h->dev_opts_host.vec_rows ? s.nit + 1 : h->dev_opts_host.vec_rows;
    ZF_CUDA(cudaMemcpyAsync(h->h_allvecs, h->d_allvecs,
                            sizeof(double) * (size_t)rows * (size_t)h->n_cols,
                            cudaMemcpyDeviceToHost, h->st));
  }
  h->h_allvecs = nullptr;                  // one solve per zf_lasso_set_allvecs
  h->h_allvecs_rows = 0;
  ZF_CUDA(cudaStreamSynchronize(h->st));
  if (h_fun) *h_fun = s.F_x;
  if (h_nit) *h_nit = s.nit;
  if (h_status) *h_status = s.status;
  if (h_lr) *h_lr = s.lr;
  return ZF_OK;
}

/* return_all's allvecs for the NEXT device-decided solve with traces (zf_lasso_solve with
 * h_allerrs / h_allfuns, or zf_lasso_dev_begin(want_trace = 1)): rows x n_cols doubles on the
 * host, row k = x^k, rows 0 .. min(nit, rows - 1) are written (proximal_gradient.py:471, 522). */
extern "C" int zf_lasso_set_allvecs(zf_lasso* h, double* h_allvecs, int64_t rows) {
  if (!h) return zf::zf_fail(ZF_ERR_INVALID, "NULL handle");
  h->h_allvecs = rows > 0 ? h_allvecs : nullptr;
  h->h_allvecs_rows = h_allvecs ? rows : 0;
  return ZF_OK;
}

// whole solve on one GPU with the device-decided loop: chunks of slots are enqueued one chunk
// ahead of the poll that looks at the previous one, so the GPU never waits for the host
static int lasso_solve_dev(zf_lasso* h, const zf_options* opt, const double* d_x0, double* d_x,
                           double* h_fun, int64_t* h_nit, int32_t* h_status, double* h_allerrs,
                           double* h_allfuns) {
  // the caller's stream may be the legacy default stream, which cannot be captured into a
  // graph: run the loop on the handle's own stream, ordered after / before the caller's
  cudaStream_t user = h->st;
  if (!h->st_own) {
    ZF_CUDA(cudaStreamCreateWithFlags(&h->st_own, cudaStreamNonBlocking));
    ZF_CUDA(cudaEventCreateWithFlags(&h->ev_own, cudaEventDisableTiming));
  }
  ZF_CUDA(cudaEventRecord(h->ev_own, user));
  ZF_CUDA(cudaStreamWaitEvent(h->st_own, h->ev_own, 0));
  h->st = h->st_own;
  int rc = zf_lasso_dev_begin(h, opt, d_x0, 0, (h_allerrs || h_allfuns) ? 1 : 0);
  if (rc == ZF_OK) rc = dev_stage(h, DS_INIT);
  const int chunk = DEV_GRAPH_SLOTS;
  int cur = 0;
  int32_t done = 0;
  if (rc == ZF_OK) rc = zf_lasso_dev_slots(h, chunk);
  if (rc == ZF_OK) rc = dev_snapshot(h, cur);
  while (rc == ZF_OK) {
    rc = zf_lasso_dev_slots(h, chunk);
    if (rc == ZF_OK) rc = dev_snapshot(h, 1 - cur);
    if (rc == ZF_OK) rc = zf_lasso_dev_poll(h, cur, 1, &done, nullptr);
    if (rc != ZF_OK || done) break;
    cur = 1 - cur;
  }
  if (rc == ZF_OK && h->dev_fixed) rc = dev_stage(h, DS_FEVAL_FINAL);
  if (rc == ZF_OK) rc = dev_stage(h, DS_FINAL);
  if (rc == ZF_OK) rc = zf_lasso_dev_finish(h, d_x, h_fun, h_nit, h_status, nullptr, h_allerrs, h_allfuns);
  h->dev_active = false;
  h->st = user;
  if (rc == ZF_OK) {
    ZF_CUDA(cudaEventRecord(h->ev_own, h->st_own));
    ZF_CUDA(cudaStreamWaitEvent(user, h->ev_own, 0));
  }
  return rc;
}


/* ---- row-sharded runs on one node: exchange through peer memory instead of NCCL ----------
 * zf_lasso_p2p_export: allocate this rank's exchange buffer and write its 64-byte
 * cudaIpcMemHandle_t to `handle_out`.  zf_lasso_p2p_attach: `handles` = the world x 64 bytes of
 * every rank's export (all-gathered by the caller over its control plane); maps the peers'
 * buffers (NVLink peer access).  Afterwards a sharded device-decided run needs no all-reduce
 * between its stages (zf_lasso_p2p_active() == 1).                                        */
extern "C" int zf_lasso_p2p_export(zf_lasso* h, void* handle_out) {
  if (!h || !handle_out) return zf::zf_fail(ZF_ERR_INVALID, "NULL argument");
  if (!h->p2p_mine) {
    const long long stride = ((h->n_cols + 1 + 31) / 32) * 32;
    const size_t bytes = sizeof(zf::P2PBuf) + sizeof(double) * 2 * (size_t)stride;
    ZF_CUDA(cudaMalloc((void**)&h->p2p_mine, bytes));
    ZF_CUDA(cudaMemset(h->p2p_mine, 0, bytes));
    ZF_CUDA(cudaMalloc((void**)&h->ss_gather, sizeof(double) * zf::P2P_MAX_RANKS));
    ZF_CUDA(cudaMalloc((void**)&h->p2p_ticket, sizeof(unsigned int)));
    ZF_CUDA(cudaMemset(h->p2p_ticket, 0, sizeof(unsigned int)));
    h->pp.stride = stride;
  }
  cudaIpcMemHandle_t hd;
  ZF_CUDA(cudaIpcGetMemHandle(&hd, h->p2p_mine));
  std::memcpy(handle_out, &hd, sizeof(hd));
  return ZF_OK;
}

extern "C" int zf_lasso_p2p_attach(zf_lasso* h, int32_t rank, int32_t world, const void* handles) {
  if (h && world == 0) {           // the ranks did not all succeed: back to the caller's all-reduce
    h->p2p_on = false;
    return ZF_OK;
  }
  if (!h || !handles || !h->p2p_mine) return zf::zf_fail(ZF_ERR_INVALID, "zf_lasso_p2p_attach before export");
  if (world < 1 || world > zf::P2P_MAX_RANKS || rank < 0 || rank >= world)
    return zf::zf_fail(ZF_ERR_UNSUPPORTED, "peer exchange supports 1..%d ranks of one node", zf::P2P_MAX_RANKS);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  for (int r = 0; r < world; ++r) {
    if (r == rank) { h->pp.buf[r] = h->p2p_mine; continue; }
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, static_cast<const char*>(handles) + 64 * (size_t)r, sizeof(hd));
    void* ptr = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return zf::zf_fail_cuda(e, "cudaIpcOpenMemHandle (peer exchange unavailable)");
    }
    h->p2p_opened[r] = ptr;
    h->pp.buf[r] = static_cast<zf::P2PBuf*>(ptr);
  }
  h->pp.world = world;
  h->pp.rank = rank;
  h->p2p_on = true;
  return ZF_OK;
}

extern "C" int zf_lasso_p2p_active(zf_lasso* h) { return (h && h->p2p_on) ? 1 : 0; }
