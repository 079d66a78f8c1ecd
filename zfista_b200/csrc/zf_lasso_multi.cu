// Dense LASSO with MANY runs sharing one A (BASELINE north_star (c): "dense A uses FP64
// tensor-core DGEMM only when many right-hand sides share A").
//
//   run k:  minimise  scale*||A x - b_k||^2 + l1*||x||_1   from x0_k with momentum (a_k, b_k)
//
// The reference would call minimize_proximal_gradient once per run (joblib fan-out over
// (a, b) pairs / starting points / observations, examples/PGM_experiment_with_various_a_b.ipynb
// and examples/cameraman.ipynb) with the dense closures of tests/test_proximal_gradient.py:49-63,
// so every run streams A through the CPU caches twice per iteration.  Here the K <= 32 runs of a
// call advance in lockstep and one gradient evaluation of ALL runs is two skinny DGEMMs
//
//   pass 1   R = A  V - B      (n_rows x K)    V = [y_1 .. y_K]
//   pass 2   G = A^T R         (n_cols x K)
//
// each of which reads A from HBM exactly once, whatever K is: the HBM cost per run drops from
// one (fused) pass to 2/K passes.  Both passes are hand-written for sm_100a:
//   * a producer thread streams 64x64 tiles of A (and the matching slice of V or R) into a ring
//     of shared-memory stages with 2-D tensor-map TMA copies (cp.async.bulk.tensor.2d +
//     mbarrier complete_tx; four 64-row x 128-byte boxes per tile, SWIZZLE_128B; out-of-range
//     rows / columns arrive as zeros).  The first version issued one 512-byte 1-D bulk copy
//     per tile row and ran at 0.3 of the HBM roofline whatever K was: ~60 cycles per copy
//     instruction, 80 of them per stage;
//   * eight consumer warps feed the FP64 tensor cores (mma.sync m8n8k4 f64 = SASS DMMA.8x8x4)
//     straight from shared memory; with the 128-byte swizzle (and the rows of an m- / n-tile
//     taken in the order 0,5,2,7,4,1,6,3) every 16-byte fragment load is bank-conflict free
//     in both passes;
//   * stages are handed back through a second set of mbarriers, so the copy of tile k+NS
//     overlaps the tensor-core work on tile k.
// Runs are independent (own step size, momentum, line search, stop test): the host keeps the
// reference's scalar logic per run (proximal_gradient.py:474-555) and a finished run simply
// stops being updated, so each run's iterates are what a solo solve produces.
#include <cuda.h>

#include <cmath>
#include <cstring>
#include <new>

#include "zf_common.cuh"
#include "zf_host.h"

namespace zf {
namespace multi {

constexpr int MAX_RUNS = 32;
constexpr int TILE = 64;        // rows and columns of one A stage
constexpr int CONSUMERS = 8;    // tensor-core warps
constexpr int THREADS = (CONSUMERS + 1) * 32;

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
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// one box of a 2-D tensor map: c0 = coordinate along the contiguous dimension, c1 = row
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1,
                                            unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1),
        "r"(smem_u32(bar)) : "memory");
}
// bounded wait: a protocol bug must trap, not hang the GPU
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned done = 0;
  for (unsigned spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (!done && spin > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void consumer_barrier() {
  asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS * 32) : "memory");
}
// D(8x8) += A(8x4, row) * B(4x8, col), fp64 tensor core.  Lane l = 4g + q holds
// A[g][q], B[q][g] and D[g][2q], D[g][2q+1].
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

constexpr int BOXC = 16;        // doubles per box row: 128 bytes, the swizzle span
constexpr int NBOX = TILE / BOXC;

template <int NT>
struct Cfg {
  static constexpr int KP = 8 * NT;               // runs, padded to whole n-tiles
  // one stage = A tile (4 boxes of 64 rows x 128 B) + vector slice (4 boxes of KP rows x 128 B)
  static constexpr int BOX_A = TILE * BOXC;       // doubles
  static constexpr int BOX_V = KP * BOXC;
  static constexpr int STAGE = NBOX * (BOX_A + BOX_V);
  static constexpr int NS1 = 4;
  static constexpr int NS2 = (NT <= 3) ? 4 : 3;
  static constexpr size_t P1_SMEM = 1024 + sizeof(double) * (size_t)(NS1 * STAGE + KP * TILE) +
                                    sizeof(unsigned long long) * 2 * NS1;
  static constexpr size_t P2_SMEM = 1024 + sizeof(double) * (size_t)(NS2 * STAGE + 2 * KP * TILE) +
                                    sizeof(unsigned long long) * 2 * NS2;
};

// the swizzle pattern repeats every 1024 bytes: boxes must start on that boundary
__device__ __forceinline__ double* align_1024(unsigned char* raw) {
  const unsigned a = smem_u32(raw);
  return reinterpret_cast<double*>(raw + ((1024u - (a & 1023u)) & 1023u));
}
// Row order inside an 8-row m- / n-tile: tensor-core row g reads physical row g ^ 4*(g & 1)
// (0,5,2,7,4,1,6,3).  Under SWIZZLE_128B the 16-byte chunk c of physical row r sits at chunk
// c ^ (r & 7); a quarter warp (g = 2j, 2j+1; q = 0..3) then touches chunks whose bit 2
// differs between its two rows, i.e. all 32 banks once.
__device__ __forceinline__ int tile_row(int g) { return g ^ ((g & 1) << 2); }

// ---------------------------------------------------------------------------------------
// pass 1:  R[k][i] = sum_j A[i][j] * V_k[j] - b_k[i],  block partials of sum_i R[k][i]^2
// Persistent CTAs over 64-row blocks; the ring streams the block's 64-column tiles.  Consumer
// warp w owns columns 8w..8w+7 of every tile (two k-steps of the m8n8k4 shape, taken from one
// 16-byte load: a lane's .x feeds the even column, .y the odd one, in A and V alike) for all
// eight 8-row m-tiles and all NT run tiles; the eight partial products are added through
// shared memory once per row block, in warp order (fixed summation order).
// ---------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(THREADS, 1)
multi_residual_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmV,
                      long long n_rows, long long n_cols, const double* __restrict__ b,
                      long long b_stride, double* __restrict__ R, long long pitch_r,
                      double* __restrict__ sq_part, const int* __restrict__ skip) {
  if (skip != nullptr && *reinterpret_cast<const volatile int*>(skip) != 0) return;   // device-decided loop
  using C_ = Cfg<NT>;
  constexpr int KP = C_::KP, NS = C_::NS1, BOX_A = C_::BOX_A, BOX_V = C_::BOX_V, STAGE = C_::STAGE;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* stages = align_1024(smem_raw);
  double* red = stages + NS * STAGE;                         // [KP][TILE]
  unsigned long long* full = reinterpret_cast<unsigned long long*>(red + KP * TILE);
  unsigned long long* empty = full + NS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], CONSUMERS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long n_rb = (n_rows + TILE - 1) / TILE;
  const int n_ch = (int)((n_cols + TILE - 1) / TILE);
  const int tail_cols = (int)(n_cols - (long long)(n_ch - 1) * TILE);    // 2..64, even

  if (warp == CONSUMERS) {
    // ------------------------------------------------------------------ producer thread
    if (lane != 0) return;
    long long it = 0;
    for (long long rb = blockIdx.x; rb < n_rb; rb += gridDim.x) {
      const int row0 = (int)(rb * TILE);
      for (int c = 0; c < n_ch; ++c, ++it) {
        const int s = (int)(it % NS);
        if (it >= NS) mbar_wait(&empty[s], (unsigned)((it / NS - 1) & 1));
        // boxes wholly past n_cols are not fetched (their lanes multiply zeros, see `dead`)
        const int nbox = (c == n_ch - 1) ? (tail_cols + BOXC - 1) / BOXC : NBOX;
        mbar_expect_tx(&full[s], (unsigned)(nbox * (BOX_A + BOX_V) * sizeof(double)));
        double* sa = stages + (size_t)s * STAGE;
        double* sv = sa + NBOX * BOX_A;
        for (int bx = 0; bx < nbox; ++bx) {
          tma_load_2d(sa + bx * BOX_A, &tmA, c * TILE + bx * BOXC, row0, &full[s]);
          tma_load_2d(sv + bx * BOX_V, &tmV, c * TILE + bx * BOXC, 0, &full[s]);
        }
      }
    }
    return;
  }
  // -------------------------------------------------------------------- consumer warps
  const int g = lane >> 2, q = lane & 3;
  const int rho = tile_row(g);
  const int col = warp * 8 + 2 * q;            // this lane's column pair inside a tile
  const int chunk = 4 * (warp & 1) + q;        // its 16-byte chunk inside the box row
  const int a_off = (warp >> 1) * BOX_A + rho * BOXC + ((chunk ^ rho) << 1);
  const int v_off = NBOX * BOX_A + (warp >> 1) * BOX_V + rho * BOXC + ((chunk ^ rho) << 1);
  double ssacc[2 * NT];
#pragma unroll
  for (int i = 0; i < 2 * NT; ++i) ssacc[i] = 0.0;
  long long it = 0;
  for (long long rb = blockIdx.x; rb < n_rb; rb += gridDim.x) {
    const long long row0 = rb * TILE;
    double acc[8][NT][2];
#pragma unroll
    for (int mt = 0; mt < 8; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
    for (int c = 0; c < n_ch; ++c, ++it) {
      const int s = (int)(it % NS);
      mbar_wait(&full[s], (unsigned)((it / NS) & 1));
      const double* sa = stages + (size_t)s * STAGE + a_off;
      const double* sv = stages + (size_t)s * STAGE + v_off;
      // columns past n_cols in the last tile must contribute 0 (TMA zero-fills a partial box,
      // a box wholly out of range is skipped and its stage memory is stale)
      const bool dead = (c == n_ch - 1) && (col >= tail_cols);
      double2 vb[NT];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        vb[nt] = *reinterpret_cast<const double2*>(sv + nt * 8 * BOXC);
        if (dead) vb[nt] = make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int mt = 0; mt < 8; ++mt) {
        double2 a = *reinterpret_cast<const double2*>(sa + mt * 8 * BOXC);
        if (dead) a = make_double2(0.0, 0.0);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          dmma(acc[mt][nt][0], acc[mt][nt][1], a.x, vb[nt].x);
          dmma(acc[mt][nt][0], acc[mt][nt][1], a.y, vb[nt].y);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
    // add the eight warps' partial products in warp order: red[run][row]
#pragma unroll 1
    for (int w = 0; w < CONSUMERS; ++w) {
      if (warp == w) {
#pragma unroll
        for (int mt = 0; mt < 8; ++mt)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int idx = (8 * nt + tile_row(2 * q + e)) * TILE + 8 * mt + rho;
              red[idx] = (w == 0) ? acc[mt][nt][e] : red[idx] + acc[mt][nt][e];
            }
      }
      consumer_barrier();
    }
    {
      const int row = tid & (TILE - 1);
      const long long grow = row0 + row;
      if (grow < n_rows) {
#pragma unroll
        for (int i = 0; i < 2 * NT; ++i) {
          const int k = (tid >> 6) + 4 * i;
          const double val = red[k * TILE + row] - b[(long long)k * b_stride + grow];
          R[(long long)k * pitch_r + grow] = val;
          ssacc[i] += val * val;
        }
      }
    }
    consumer_barrier();                          // red is free for the next row block
  }
  // sum r^2 per run: rows live in two warps per run quarter (tid >> 6), fixed order
#pragma unroll
  for (int i = 0; i < 2 * NT; ++i) ssacc[i] = warp_sum(ssacc[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 2 * NT; ++i) red[warp * 2 * NT + i] = ssacc[i];
  }
  consumer_barrier();
  if (tid < KP) {
    const int j = tid & 3, i = tid >> 2;         // run = j + 4 i, held by warps 2j and 2j+1
    sq_part[(long long)blockIdx.x * KP + tid] =
        red[(2 * j) * 2 * NT + i] + red[(2 * j + 1) * 2 * NT + i];
  }
}

// ---------------------------------------------------------------------------------------
// pass 2:  gpart[split][k][j] = sum_{i in split} A[i][j] * R[k][i]
// Work item = (64-column slab, row split); persistent CTAs take items round-robin.  The ring
// streams 64-row tiles of the slab and the matching 64 residuals of every run.  A^T is the
// tensor-core "A" operand: consumer warp w = (cg = w & 3, kh = w >> 2) owns columns
// 16cg..16cg+15 (even columns = one m-tile, odd columns = the other, from one 16-byte load)
// and rows 32kh..32kh+31 of the tile; the two row halves are added through shared memory.
// ---------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(THREADS, 1)
multi_atr_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmR,
                 long long n_rows, long long n_cols, long long rows_per_split, int n_splits,
                 double* __restrict__ gpart, long long pitch_c, const int* __restrict__ skip) {
  if (skip != nullptr && *reinterpret_cast<const volatile int*>(skip) != 0) return;   // device-decided loop
  using C_ = Cfg<NT>;
  constexpr int KP = C_::KP, NS = C_::NS2, BOX_A = C_::BOX_A, BOX_R = C_::BOX_V, STAGE = C_::STAGE;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* stages = align_1024(smem_raw);
  double* red = stages + NS * STAGE;                         // [2][KP][TILE]
  unsigned long long* full = reinterpret_cast<unsigned long long*>(red + 2 * KP * TILE);
  unsigned long long* empty = full + NS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], CONSUMERS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long n_slabs = (n_cols + TILE - 1) / TILE;
  const long long n_items = n_slabs * n_splits;

  if (warp == CONSUMERS) {
    if (lane != 0) return;
    long long it = 0;
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
      const long long slab = item % n_slabs, split = item / n_slabs;
      const int col0 = (int)(slab * TILE);
      const int seg = (int)((n_cols - col0 < TILE) ? n_cols - col0 : TILE);
      const int nbox_a = (seg + BOXC - 1) / BOXC;    // columns past n_cols are never stored
      const long long r_lo = split * rows_per_split;
      const long long r_hi = (r_lo + rows_per_split < n_rows) ? r_lo + rows_per_split : n_rows;
      for (long long rc = r_lo; rc < r_hi; rc += TILE, ++it) {
        const int s = (int)(it % NS);
        if (it >= NS) mbar_wait(&empty[s], (unsigned)((it / NS - 1) & 1));
        const int valid = (int)((r_hi - rc < TILE) ? r_hi - rc : TILE);
        const int nbox_r = (valid + BOXC - 1) / BOXC;
        mbar_expect_tx(&full[s], (unsigned)((nbox_a * BOX_A + nbox_r * BOX_R) * sizeof(double)));
        double* sa = stages + (size_t)s * STAGE;
        double* sr = sa + NBOX * BOX_A;
        for (int bx = 0; bx < nbox_a; ++bx)
          tma_load_2d(sa + bx * BOX_A, &tmA, col0 + bx * BOXC, (int)rc, &full[s]);
        for (int bx = 0; bx < nbox_r; ++bx)
          tma_load_2d(sr + bx * BOX_R, &tmR, (int)rc + bx * BOXC, 0, &full[s]);
      }
    }
    return;
  }
  const int g = lane >> 2, q = lane & 3;
  const int rho = tile_row(g);
  const int cg = warp & 3, kh = warp >> 2;
  long long it = 0;
  for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
    const long long slab = item % n_slabs, split = item / n_slabs;
    const long long col0 = slab * TILE;
    const int seg = (int)((n_cols - col0 < TILE) ? n_cols - col0 : TILE);
    const long long r_lo = split * rows_per_split;
    const long long r_hi = (r_lo + rows_per_split < n_rows) ? r_lo + rows_per_split : n_rows;
    double acc[2][NT][2];
#pragma unroll
    for (int e = 0; e < 2; ++e)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[e][nt][0] = acc[e][nt][1] = 0.0;
    for (long long rc = r_lo; rc < r_hi; rc += TILE, ++it) {
      const int s = (int)(it % NS);
      mbar_wait(&full[s], (unsigned)((it / NS) & 1));
      const int valid = (int)((r_hi - rc < TILE) ? r_hi - rc : TILE);
      const double* sa = stages + (size_t)s * STAGE + cg * BOX_A;
      const double* sr = stages + (size_t)s * STAGE + NBOX * BOX_A + rho * BOXC;
#pragma unroll
      for (int sp = 0; sp < 4; ++sp) {
        const int rbase = 32 * kh + 8 * sp + 2 * q;     // this lane's row pair (k index)
        // rows rbase, rbase+1 of the tile: (row & 7) = 2q, 2q+1
        double2 a0 = *reinterpret_cast<const double2*>(sa + rbase * BOXC + ((g ^ (2 * q)) << 1));
        double2 a1 = *reinterpret_cast<const double2*>(sa + (rbase + 1) * BOXC +
                                                       ((g ^ (2 * q + 1)) << 1));
        const int chunk = 4 * (sp & 1) + q;              // residual pair inside its 16-row box
        const double* srb = sr + (2 * kh + (sp >> 1)) * BOX_R + ((chunk ^ rho) << 1);
        double2 rv[NT];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
          rv[nt] = *reinterpret_cast<const double2*>(srb + nt * 8 * BOXC);
        if (valid < TILE) {                              // rows past the split: contribute 0
          if (rbase >= valid) {
            a0 = make_double2(0.0, 0.0);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) rv[nt].x = 0.0;
          }
          if (rbase + 1 >= valid) {
            a1 = make_double2(0.0, 0.0);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) rv[nt].y = 0.0;
          }
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          dmma(acc[0][nt][0], acc[0][nt][1], a0.x, rv[nt].x);
          dmma(acc[1][nt][0], acc[1][nt][1], a0.y, rv[nt].x);
          dmma(acc[0][nt][0], acc[0][nt][1], a1.x, rv[nt].y);
          dmma(acc[1][nt][0], acc[1][nt][1], a1.y, rv[nt].y);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
    // red[kh][run][col] ; then (kh = 0) + (kh = 1), coalesced store of the slab's partial
#pragma unroll
    for (int e = 0; e < 2; ++e)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int j = 0; j < 2; ++j)
          red[(kh * KP + 8 * nt + tile_row(2 * q + j)) * TILE + 16 * cg + 2 * g + e] = acc[e][nt][j];
    consumer_barrier();
    for (int e = tid; e < KP * TILE; e += CONSUMERS * 32) {
      const int k = e >> 6, cc = e & (TILE - 1);
      if (cc < seg)
        gpart[((long long)split * KP + k) * pitch_c + col0 + cc] = red[e] + red[KP * TILE + e];
    }
    consumer_barrier();
  }
}

// partial[k][j] = sum_split gpart[split][k][j] (fixed order); ss[k] = sum_blk sq_part[blk][k]
// `partial` = [KP][pitch_c] followed by ss[KP]: the one buffer a row-sharded run all-reduces.
__global__ void __launch_bounds__(256)
multi_collect_kernel(const double* __restrict__ gpart, int n_splits, const double* __restrict__ sq_part,
                     int n_sq, int kp, long long pitch_c, long long n_cols,
                     double* __restrict__ partial) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (gpart && j < n_cols) {
    double s = 0.0;
    for (int sp = 0; sp < n_splits; ++sp) s += gpart[((long long)sp * kp + k) * pitch_c + j];
    partial[(long long)k * pitch_c + j] = s;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double s = 0.0;
    for (int bl = 0; bl < n_sq; ++bl) s += sq_part[(long long)bl * kp + k];
    partial[(long long)kp * pitch_c + k] = s;
  }
}

constexpr int VEC_THREADS = 256;
constexpr int VEC_BLOCKS = 64;     // per run

struct StepSums { double gd, dd, abs1, maxd; };

struct StepArgs {
  double lr[MAX_RUNS];       // prox: step size;  momentum: (t_k - 1) / t_{k+1}
  unsigned mask;             // runs this launch touches
};

// per run k in `mask`:  x_new = soft(y - lr*g, lr*l1) with g = src[k]*two_scale
// (tests/test_proximal_gradient.py:55-63), g stored to g_out (first trial of an iteration), and
//   gd = g.(x-y), dd = ||x-y||^2, abs1 = ||x||_1, maxd = max|x-y|
// abs_only: only abs1 of the previous iterate (g(x0) at start-up).
__global__ void __launch_bounds__(VEC_THREADS)
multi_prox_kernel(StepArgs a, const double* __restrict__ Y, const double* __restrict__ Xp,
                  double* __restrict__ Xn, long long pitch_c, const double* __restrict__ src,
                  double two_scale, double l1, long long n, double* __restrict__ g_out,
                  StepSums* __restrict__ block_sums, unsigned int* __restrict__ counter,
                  StepSums* __restrict__ out, int abs_only) {
  const int k = blockIdx.y;
  if (!((a.mask >> k) & 1u)) return;
  __shared__ StepSums sh[VEC_THREADS / 32];
  __shared__ bool is_last;
  const double* xp = Xp + (long long)k * pitch_c;
  double* xn = Xn + (long long)k * pitch_c;
  const double* y = Y + (long long)k * pitch_c;
  const double* gs = src + (long long)k * pitch_c;
  double* go = g_out ? g_out + (long long)k * pitch_c : nullptr;
  // (thresh is the ROUNDED product, as numpy forms l1 * lr; left to nvcc it may or may not be
  // contracted into the subtraction inside soft_threshold, differently in different kernels)
  const double lr = a.lr[k], thresh = __dmul_rn(lr, l1);
  StepSums s{0.0, 0.0, 0.0, 0.0};
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n;
       j += (long long)gridDim.x * blockDim.x) {
    if (abs_only) {
      s.abs1 += fabs(xp[j]);
    } else {
      const double gj = gs[j] * two_scale;
      const double yj = y[j];
      const double xj = soft_threshold(fma(-lr, gj, yj), thresh);
      const double d = xj - yj;
      xn[j] = xj;
      if (go) go[j] = gj;
      s.gd = fma(gj, d, s.gd);
      s.dd = fma(d, d, s.dd);
      s.abs1 += fabs(xj);
      s.maxd = fmax(s.maxd, fabs(d));
    }
  }
  s.gd = warp_sum(s.gd);
  s.dd = warp_sum(s.dd);
  s.abs1 = warp_sum(s.abs1);
  s.maxd = warp_max(s.maxd);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sh[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    StepSums t = sh[0];
    for (int w = 1; w < VEC_THREADS / 32; ++w) {
      t.gd += sh[w].gd; t.dd += sh[w].dd; t.abs1 += sh[w].abs1; t.maxd = fmax(t.maxd, sh[w].maxd);
    }
    block_sums[(long long)k * gridDim.x + blockIdx.x] = t;
    __threadfence();
    const unsigned int done = atomicAdd(&counter[k], 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    const StepSums* bs = block_sums + (long long)k * gridDim.x;
    StepSums t = bs[0];
    for (unsigned int i = 1; i < gridDim.x; ++i) {
      const StepSums u = bs[i];
      t.gd += u.gd; t.dd += u.dd; t.abs1 += u.abs1; t.maxd = fmax(t.maxd, u.maxd);
    }
    out[k] = t;
    counter[k] = 0u;
  }
}

// per run k in `mask`:  y = x_new + mom*(x_new - x_prev), then x_prev = x_new
// (proximal_gradient.py:534-538).  The iterates of all runs stay in the fixed arrays Xp / Xn / Y
// (one tensor map each); a run that is not in `mask` (finished, failed) keeps all three.
__global__ void __launch_bounds__(VEC_THREADS)
multi_momentum_kernel(StepArgs a, double* __restrict__ Xp, const double* __restrict__ Xn,
                      long long pitch_c, long long n, double* __restrict__ Y) {
  const int k = blockIdx.y;
  if (!((a.mask >> k) & 1u)) return;
  double* xp = Xp + (long long)k * pitch_c;
  const double* xn = Xn + (long long)k * pitch_c;
  double* y = Y + (long long)k * pitch_c;
  const double mom = a.lr[k];
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n;
       j += (long long)gridDim.x * blockDim.x) {
    const double xj = xn[j];
    y[j] = fma(mom, xj - xp[j], xj);
    xp[j] = xj;
  }
}

// ---------------------------------------------------------------------------------------
// Device-decided rounds (as zf_lasso.cu's zf_lasso_dev_*, for K runs in lockstep): the per-run
// scalars of proximal_gradient.py:474-538 and the run masks (active / on trial / advancing) live
// in device memory, one-warp kernels take the decisions lane k for run k, the n_cols-sized
// kernels read their step sizes, momentum factors and masks from there.  A slot is
//   [gradient of all runs at Y: two DGEMM passes + collect]   (skipped while a retry is pending)
//   prox of the runs on trial -> after_prox (sums, f(y), subproblem values; fixed step: accept)
//   [residual DGEMM at the candidates + collect -> decide]    (line search / F trace only)
//   momentum of the runs that advance
// and slots enqueued after every run has ended do nothing.  Scalar arithmetic in *_rn
// intrinsics: bit for bit the host-decided loop below.
// ---------------------------------------------------------------------------------------
enum { MD_GRAD = 0, MD_RETRY = 1, MD_DONE = 2 };

struct MultiDevOpts {
  double lr0, tol, tol_internal, decay, scale, l1;
  long long max_iter, pitch_c, n;
  int max_bt, nesterov, deprecated, need_F, cap, n_runs, kp, pad;
  double* allerrs;             // [n_runs][cap]
  double* allfuns;             // [n_runs][cap + 1]
  double na[MAX_RUNS], nb[MAX_RUNS];
};

struct MultiDevRun {
  double lr, t_prev, F_prev, F_x, f_y, sub_fun, err, mom;
  StepSums sums;
  long long nit;
  int status, bt, F_known, result_is_prev;
};

struct MultiDevState {
  MultiDevRun run[MAX_RUNS];
  unsigned active, trial, accepted, adv;
  int phase, skip_grad, done, pad;
};

__device__ __forceinline__ double md_f_from_ss(double ss, double scale) {
  const double nrm = __dsqrt_rn(ss);
  return __dmul_rn(__dmul_rn(nrm, nrm), scale);
}

__device__ __forceinline__ double md_subproblem_fun(const MultiDevOpts& o, const MultiDevRun& r) {
  const double nrm = __dsqrt_rn(r.sums.dd);
  double fun = __dadd_rn(__dadd_rn(r.sums.gd, __dmul_rn(o.l1, r.sums.abs1)),
                         __ddiv_rn(__ddiv_rn(__dmul_rn(nrm, nrm), 2.0), r.lr));
  if (!o.deprecated) fun = __dadd_rn(fun, __dsub_rn(r.f_y, r.F_prev));
  return fun;
}

// one warp, lane k = run k: the runs in `accepted` have an accepted candidate -- stop tests,
// traces, t_{k+1}; then the masks and the phase of the next slot
__device__ void md_advance(const MultiDevOpts& o, MultiDevState* st, unsigned accepted, int lane) {
  bool stops = false, advances = false;
  if (lane < o.n_runs && ((accepted >> lane) & 1u)) {
    MultiDevRun& r = st->run[lane];
    r.err = r.sums.maxd;
    if (o.cap > 0 && r.nit <= o.cap) {
      if (o.allerrs) o.allerrs[(long long)lane * o.cap + r.nit - 1] = r.err;
      if (o.allfuns && r.F_known) o.allfuns[(long long)lane * (o.cap + 1) + r.nit] = r.F_x;
    }
    const bool converged = r.err < o.tol;
    if (converged || r.nit >= o.max_iter) {
      r.status = converged ? 1 : 0;
      stops = true;
    } else {
      double mom = 0.0;
      if (o.nesterov) {
        const double t = r.t_prev;
        const double tn = __dadd_rn(
            __dsqrt_rn(__dadd_rn(__dsub_rn(__dmul_rn(t, t), __dmul_rn(o.na[lane], t)), o.nb[lane])), 0.5);
        mom = __ddiv_rn(__dsub_rn(t, 1.0), tn);
        r.t_prev = tn;
      }
      r.mom = mom;
      r.F_prev = r.F_x;
      r.nit += 1;
      advances = true;
    }
  }
  const unsigned stop_mask = __ballot_sync(0xffffffffu, stops);
  const unsigned adv_mask = __ballot_sync(0xffffffffu, advances);
  if (lane == 0) {
    st->active &= ~stop_mask;
    st->adv = adv_mask;
    st->accepted = 0u;
    st->trial = 0u;
    if (st->active) {
      st->phase = MD_GRAD;
      st->skip_grad = 0;
    } else {
      st->phase = MD_DONE;
      st->skip_grad = 1;
      st->done = 1;
    }
  }
}

__global__ void multi_dev_init_kernel(const MultiDevOpts* __restrict__ op, MultiDevState* st,
                                      const StepSums* __restrict__ sums,
                                      const double* __restrict__ ss) {
  const int lane = threadIdx.x;
  if (blockIdx.x != 0 || lane >= 32) return;
  const MultiDevOpts& o = *op;
  if (lane < o.n_runs) {
    MultiDevRun& r = st->run[lane];
    const double F0 = __dadd_rn(md_f_from_ss(ss[lane], o.scale), __dmul_rn(o.l1, sums[lane].abs1));
    r.lr = o.lr0; r.t_prev = 1.0; r.F_prev = F0; r.F_x = F0; r.f_y = 0.0; r.sub_fun = 0.0;
    r.err = CUDART_INF; r.mom = 0.0; r.sums = sums[lane]; r.nit = 1; r.status = 0; r.bt = 0;
    r.F_known = 0; r.result_is_prev = 0;
    if (o.cap > 0 && o.allfuns) o.allfuns[(long long)lane * (o.cap + 1)] = F0;
  }
  if (lane == 0) {
    st->active = (o.n_runs == 32) ? 0xffffffffu : ((1u << o.n_runs) - 1u);
    st->trial = 0u; st->accepted = 0u; st->adv = 0u;
    st->phase = MD_GRAD; st->skip_grad = 0; st->done = 0;
  }
}

// prox of the runs on trial: first trial of an iteration (phase GRAD: every active run, g from
// the collected partials, stored to G) or a retry (the runs still on trial, g from G)
__global__ void __launch_bounds__(VEC_THREADS)
multi_dev_prox_kernel(const MultiDevOpts* __restrict__ op, const MultiDevState* __restrict__ st,
                      const double* __restrict__ Y, double* __restrict__ Xn,
                      const double* __restrict__ partial, double* __restrict__ G,
                      StepSums* __restrict__ block_sums, unsigned int* __restrict__ counter,
                      StepSums* __restrict__ out) {
  if (st->done != 0) return;
  const int k = blockIdx.y;
  const bool first = (st->phase == MD_GRAD);
  const unsigned mask = first ? st->active : st->trial;
  if (!((mask >> k) & 1u)) return;
  __shared__ StepSums sh[VEC_THREADS / 32];
  __shared__ bool is_last;
  const long long pitch_c = op->pitch_c, n = op->n;
  double* xn = Xn + (long long)k * pitch_c;
  const double* y = Y + (long long)k * pitch_c;
  const double* gs = (first ? partial : G) + (long long)k * pitch_c;
  double* go = G + (long long)k * pitch_c;
  const double two_scale = first ? 2.0 * op->scale : 1.0;
  const double lr = st->run[k].lr, thresh = __dmul_rn(lr, op->l1);
  StepSums s{0.0, 0.0, 0.0, 0.0};
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n;
       j += (long long)gridDim.x * blockDim.x) {
    const double gj = gs[j] * two_scale;
    const double yj = y[j];
    const double xj = soft_threshold(fma(-lr, gj, yj), thresh);
    const double d = xj - yj;
    xn[j] = xj;
    if (first) go[j] = gj;
    s.gd = fma(gj, d, s.gd);
    s.dd = fma(d, d, s.dd);
    s.abs1 += fabs(xj);
    s.maxd = fmax(s.maxd, fabs(d));
  }
  s.gd = warp_sum(s.gd);
  s.dd = warp_sum(s.dd);
  s.abs1 = warp_sum(s.abs1);
  s.maxd = warp_max(s.maxd);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sh[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    StepSums t = sh[0];
    for (int w = 1; w < VEC_THREADS / 32; ++w) {
      t.gd += sh[w].gd; t.dd += sh[w].dd; t.abs1 += sh[w].abs1; t.maxd = fmax(t.maxd, sh[w].maxd);
    }
    block_sums[(long long)k * gridDim.x + blockIdx.x] = t;
    __threadfence();
    is_last = (atomicAdd(&counter[k], 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    const StepSums* bs = block_sums + (long long)k * gridDim.x;
    StepSums t = bs[0];
    for (unsigned int i = 1; i < gridDim.x; ++i) {
      const StepSums u = bs[i];
      t.gd += u.gd; t.dd += u.dd; t.abs1 += u.abs1; t.maxd = fmax(t.maxd, u.maxd);
    }
    out[k] = t;
    counter[k] = 0u;
  }
}

// one warp after the prox: the trial runs' sums, f(y) (first trial) and subproblem values; with a
// fixed step and no F trace every trial is accepted here
__global__ void multi_dev_after_prox_kernel(const MultiDevOpts* __restrict__ op, MultiDevState* st,
                                            const StepSums* __restrict__ sums,
                                            const double* __restrict__ ss) {
  const int lane = threadIdx.x;
  if (blockIdx.x != 0 || lane >= 32 || st->done != 0) return;
  const MultiDevOpts& o = *op;
  const bool first = (st->phase == MD_GRAD);
  const unsigned mask = first ? st->active : st->trial;
  if (lane < o.n_runs && ((mask >> lane) & 1u)) {
    MultiDevRun& r = st->run[lane];
    r.sums = sums[lane];
    if (first) {
      r.bt = 0;
      r.f_y = md_f_from_ss(ss[lane], o.scale);
      r.F_known = 0;
    }
    r.sub_fun = md_subproblem_fun(o, r);
  }
  __syncwarp();
  if (!o.need_F) {
    md_advance(o, st, mask, lane);
  } else if (lane == 0) {
    st->trial = mask;
    st->adv = 0u;
  }
}

// one warp after the residual pass at the candidates: F(x_new) and the line-search test of every
// run on trial; accepted runs wait until no run is on trial any more, then all advance together
__global__ void multi_dev_decide_kernel(const MultiDevOpts* __restrict__ op, MultiDevState* st,
                                        const double* __restrict__ ss) {
  const int lane = threadIdx.x;
  if (blockIdx.x != 0 || lane >= 32 || st->done != 0) return;
  const MultiDevOpts& o = *op;
  const unsigned trial = st->trial;
  bool acc = false, fail = false;
  if (lane < o.n_runs && ((trial >> lane) & 1u)) {
    MultiDevRun& r = st->run[lane];
    const double f_x = md_f_from_ss(ss[lane], o.scale);
    r.F_x = __dadd_rn(f_x, __dmul_rn(o.l1, r.sums.abs1));
    r.F_known = 1;
    bool ok;
    if (o.decay == 1.0) ok = true;                                        // proximal_gradient.py:298
    else if (o.deprecated) ok = (__dsub_rn(f_x, r.f_y) <= __dadd_rn(r.sub_fun, o.tol_internal));
    else ok = (__dsub_rn(r.F_x, r.F_prev) <= __dadd_rn(r.sub_fun, o.tol_internal));
    if (ok) {
      acc = true;
    } else {
      r.lr = __dmul_rn(r.lr, o.decay);
      r.bt += 1;
      if (r.bt >= o.max_bt) {
        // RuntimeError("Backtracking failed ...") -> x = x_prev, nit - 1 (proximal_gradient.py:493-509)
        r.result_is_prev = 1;
        r.F_x = r.F_prev;
        r.nit -= 1;
        r.status = -1;
        fail = true;
      }
    }
  }
  const unsigned acc_mask = __ballot_sync(0xffffffffu, acc);
  const unsigned fail_mask = __ballot_sync(0xffffffffu, fail);
  const unsigned left = trial & ~(acc_mask | fail_mask);
  const unsigned accepted = st->accepted | acc_mask;
  __syncwarp();
  if (lane == 0) {
    st->active &= ~fail_mask;
    st->accepted = accepted;
    st->trial = left;
  }
  __syncwarp();
  if (left) {
    if (lane == 0) {
      st->phase = MD_RETRY;          // same gradients, smaller steps for the runs still on trial
      st->skip_grad = 1;
      st->adv = 0u;
    }
  } else {
    md_advance(o, st, accepted, lane);
  }
}

// y = x_new + mom (x_new - x_prev), x_prev = x_new for the runs that advance
__global__ void __launch_bounds__(VEC_THREADS)
multi_dev_momentum_kernel(const MultiDevOpts* __restrict__ op, const MultiDevState* __restrict__ st,
                          double* __restrict__ Xp, const double* __restrict__ Xn,
                          double* __restrict__ Y) {
  const int k = blockIdx.y;
  if (!((st->adv >> k) & 1u)) return;
  const long long pitch_c = op->pitch_c, n = op->n;
  double* xp = Xp + (long long)k * pitch_c;
  const double* xn = Xn + (long long)k * pitch_c;
  double* y = Y + (long long)k * pitch_c;
  const double mom = st->run[k].mom;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n;
       j += (long long)gridDim.x * blockDim.x) {
    const double xj = xn[j];
    y[j] = fma(mom, xj - xp[j], xj);
    xp[j] = xj;
  }
}

// res.fun of the runs that never evaluated F at their result
__global__ void multi_dev_final_kernel(const MultiDevOpts* __restrict__ op, MultiDevState* st,
                                       const double* __restrict__ ss) {
  const int lane = threadIdx.x;
  if (blockIdx.x != 0 || lane >= op->n_runs) return;
  MultiDevRun& r = st->run[lane];
  if (r.F_known || r.result_is_prev) return;
  r.F_x = __dadd_rn(md_f_from_ss(ss[lane], op->scale), __dmul_rn(op->l1, r.sums.abs1));
  r.F_known = 1;
}

// grad[k][j] = partial[k][j] * 2*scale ; f[k] = ||r_k||^2 * scale   (bench / closures)
__global__ void __launch_bounds__(VEC_THREADS)
multi_scale_kernel(const double* __restrict__ partial, int kp, long long pitch_c, long long n,
                   double two_scale, double scale, double* __restrict__ grad,
                   double* __restrict__ f_out) {
  const int k = blockIdx.y;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n;
       j += (long long)gridDim.x * blockDim.x)
    grad[(long long)k * n + j] = partial[(long long)k * pitch_c + j] * two_scale;
  if (f_out && blockIdx.x == 0 && threadIdx.x == 0) {
    const double nrm = sqrt(partial[(long long)kp * pitch_c + k]);
    f_out[k] = nrm * nrm * scale;
  }
}

}  // namespace multi
}  // namespace zf

// =======================================================================================
// handle
// =======================================================================================
namespace {
enum MultiPhase { MP_IDLE = 0, MP_INIT, MP_GRAD, MP_FNEW, MP_FINAL, MP_DONE };

struct RunState {
  double lr = 1.0, t_prev = 1.0, F_prev = 0.0, F_x = 0.0, f_y = 0.0, sub_fun = 0.0, err = 0.0;
  double a = 0.0, b = 0.25;
  long long nit = 0;
  int status = 0, bt = 0;
  bool F_known = false, result_is_prev = false;
  zf::multi::StepSums sums{};
};
}  // namespace

struct zf_lasso_multi {
  const double* A = nullptr;
  long long n_rows = 0, n_cols = 0, pitch_c = 0, pitch_r = 0;
  int n_runs = 0, nt = 1, kp = 8;
  double scale = 1.0, l1 = 0.0;
  cudaStream_t st = 0;
  int n_sm = 148, grid = 148, n_splits = 1;
  long long rows_per_split = 0;
  bool b_batched = false;
  // device workspace
  double* vecs = nullptr;      // 4 * kp * pitch_c : Xp, Xn, Y, G (rows of padding runs stay 0)
  double *Xp = nullptr, *Xn = nullptr, *Y = nullptr, *G = nullptr;
  CUtensorMap tmA, tmY, tmXn, tmR;   // SWIZZLE_128B boxes of 16 doubles x (64 | kp) rows
  double* bcopy = nullptr;     // kp * pitch_r (batched b) or pitch_r
  double* R = nullptr;         // kp * pitch_r
  double* gpart = nullptr;     // n_splits * kp * pitch_c
  double* sq_part = nullptr;   // grid * kp
  double* partial = nullptr;   // kp * pitch_c + kp
  zf::multi::StepSums* block_sums = nullptr;   // kp * VEC_BLOCKS
  zf::multi::StepSums* d_sums = nullptr;       // kp
  unsigned int* counter = nullptr;             // kp
  double* h_pin = nullptr;                     // pinned: StepSums[kp] | ss[kp]
  // solver state
  zf_options opt{};
  int phase = MP_IDLE;
  bool need_F = false;
  unsigned active = 0, trial = 0, accepted = 0;
  RunState run[zf::multi::MAX_RUNS];
  double* h_allerrs = nullptr;
  double* h_allfuns = nullptr;
  // device-decided rounds
  const int* skip = nullptr;                     // flag the DGEMM passes test
  zf::multi::MultiDevOpts* d_opts = nullptr;
  zf::multi::MultiDevState* d_state = nullptr;
  zf::multi::MultiDevState* h_state = nullptr;   // pinned, 2 snapshots
  cudaEvent_t ev_poll[2] = {nullptr, nullptr};
  int dev_need_F = 0, dev_cap = 0;               // of the device-decided solve in flight
  double* d_allerrs = nullptr;
  double* d_allfuns = nullptr;
  size_t dev_trace = 0;
};

namespace {

#define ZF_CUDA(call)                                          \
  do {                                                         \
    cudaError_t _e = (call);                                   \
    if (_e != cudaSuccess) return zf::zf_fail_cuda(_e, #call); \
  } while (0)

using zf::multi::Cfg;
using zf::multi::StepArgs;
using zf::multi::StepSums;

// which = 0: the vectors are the y_k, 1: the candidates x_new_k
template <int NT>
int launch_residual_t(zf_lasso_multi* h, int which) {
  auto k = zf::multi::multi_residual_kernel<NT>;
  ZF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)Cfg<NT>::P1_SMEM));
  k<<<h->grid, zf::multi::THREADS, Cfg<NT>::P1_SMEM, h->st>>>(
      h->tmA, which == 0 ? h->tmY : h->tmXn, h->n_rows, h->n_cols, h->bcopy,
      h->b_batched ? h->pitch_r : 0, h->R, h->pitch_r, h->sq_part, h->skip);
  ZF_CUDA(cudaGetLastError());
  zf::zf_count_launch();
  return ZF_OK;
}

int launch_residual(zf_lasso_multi* h, int which) {
  switch (h->nt) {
    case 1: return launch_residual_t<1>(h, which);
    case 2: return launch_residual_t<2>(h, which);
    case 3: return launch_residual_t<3>(h, which);
    default: return launch_residual_t<4>(h, which);
  }
}

template <int NT>
int launch_atr_t(zf_lasso_multi* h) {
  auto k = zf::multi::multi_atr_kernel<NT>;
  ZF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)Cfg<NT>::P2_SMEM));
  k<<<h->grid, zf::multi::THREADS, Cfg<NT>::P2_SMEM, h->st>>>(
      h->tmA, h->tmR, h->n_rows, h->n_cols, h->rows_per_split, h->n_splits, h->gpart, h->pitch_c,
      h->skip);
  ZF_CUDA(cudaGetLastError());
  zf::zf_count_launch();
  return ZF_OK;
}

int launch_atr(zf_lasso_multi* h) {
  switch (h->nt) {
    case 1: return launch_atr_t<1>(h);
    case 2: return launch_atr_t<2>(h);
    case 3: return launch_atr_t<3>(h);
    default: return launch_atr_t<4>(h);
  }
}

int launch_collect(zf_lasso_multi* h, bool with_gradient) {
  dim3 grid(with_gradient ? (unsigned)((h->n_cols + 255) / 256) : 1u, (unsigned)h->kp);
  zf::multi::multi_collect_kernel<<<grid, 256, 0, h->st>>>(
      with_gradient ? h->gpart : nullptr, h->n_splits, h->sq_part, h->grid, h->kp, h->pitch_c,
      h->n_cols, h->partial);
  ZF_CUDA(cudaGetLastError());
  zf::zf_count_launch();
  return ZF_OK;
}

int gradient_pass(zf_lasso_multi* h) {
  int rc = launch_residual(h, 0);
  if (rc != ZF_OK) return rc;
  rc = launch_atr(h);
  if (rc != ZF_OK) return rc;
  return launch_collect(h, true);
}

// prox (or |x0|_1) for the runs in `mask`; their sums and every run's ss land in h_pin
int run_prox(zf_lasso_multi* h, unsigned mask, bool first_trial, bool abs_only) {
  StepArgs a{};
  for (int k = 0; k < h->n_runs; ++k) a.lr[k] = h->run[k].lr;
  a.mask = mask;
  dim3 grid(zf::multi::VEC_BLOCKS, (unsigned)h->n_runs);
  zf::multi::multi_prox_kernel<<<grid, zf::multi::VEC_THREADS, 0, h->st>>>(
      a, h->Y, h->Xp, h->Xn, h->pitch_c, first_trial ? h->partial : h->G,
      first_trial ? 2.0 * h->scale : 1.0, h->l1, h->n_cols, first_trial ? h->G : nullptr,
      h->block_sums, h->counter, h->d_sums, abs_only ? 1 : 0);
  ZF_CUDA(cudaGetLastError());
  zf::zf_count_launch();
  return ZF_OK;
}

int fetch(zf_lasso_multi* h, bool sums, unsigned sums_mask) {
  if (sums)
    ZF_CUDA(cudaMemcpyAsync(h->h_pin, h->d_sums, sizeof(StepSums) * h->kp,
                            cudaMemcpyDeviceToHost, h->st));
  ZF_CUDA(cudaMemcpyAsync(h->h_pin + 4 * h->kp, h->partial + (long long)h->kp * h->pitch_c,
                          sizeof(double) * h->kp, cudaMemcpyDeviceToHost, h->st));
  ZF_CUDA(cudaStreamSynchronize(h->st));
  if (sums)
    for (int k = 0; k < h->n_runs; ++k)
      if ((sums_mask >> k) & 1u) std::memcpy(&h->run[k].sums, h->h_pin + 4 * k, sizeof(StepSums));
  return ZF_OK;
}

inline double ss_of(const zf_lasso_multi* h, int k) { return h->h_pin[4 * h->kp + k]; }

// f = np.linalg.norm(A @ x - b) ** 2 * scale   (test_proximal_gradient.py:50)
inline double f_from_ss(const zf_lasso_multi* h, double ss) {
  const double nrm = std::sqrt(ss);
  return nrm * nrm * h->scale;
}

// proximal_gradient.py:149-155 for one objective
inline double subproblem_fun(const zf_lasso_multi* h, const RunState& r) {
  const double nrm = std::sqrt(r.sums.dd);
  double fun = r.sums.gd + h->l1 * r.sums.abs1 + nrm * nrm / 2.0 / r.lr;
  if (!h->opt.deprecated) fun += r.f_y - r.F_prev;
  return fun;
}

// every run of this round has an accepted candidate (or failed): stop tests, momentum, next
// iterates (proximal_gradient.py:510-538), run by run
int advance(zf_lasso_multi* h, int* next) {
  StepArgs m{};
  unsigned adv = 0;
  const int cap = h->opt.trace_capacity;
  for (int k = 0; k < h->n_runs; ++k) {
    if (!((h->accepted >> k) & 1u)) continue;
    RunState& r = h->run[k];
    r.err = r.sums.maxd;
    if (cap > 0 && r.nit <= cap) {
      if (h->h_allerrs) h->h_allerrs[(long long)k * cap + r.nit - 1] = r.err;
      if (h->h_allfuns && r.F_known) h->h_allfuns[(long long)k * (cap + 1) + r.nit] = r.F_x;
    }
    const bool converged = r.err < h->opt.tol;
    if (converged || r.nit >= h->opt.max_iter) {
      r.status = converged ? 1 : 0;
      h->active &= ~(1u << k);
      continue;
    }
    double mom = 0.0;
    if (h->opt.nesterov) {
      const double t = r.t_prev;
      const double t_new = std::sqrt(t * t - r.a * t + r.b) + 0.5;
      mom = (t - 1.0) / t_new;
      r.t_prev = t_new;
    }
    m.lr[k] = mom;
    adv |= 1u << k;
    r.F_prev = r.F_x;
    r.nit += 1;
  }
  h->accepted = 0;
  if (adv) {
    m.mask = adv;
    dim3 grid(zf::multi::VEC_BLOCKS, (unsigned)h->n_runs);
    zf::multi::multi_momentum_kernel<<<grid, zf::multi::VEC_THREADS, 0, h->st>>>(
        m, h->Xp, h->Xn, h->pitch_c, h->n_cols, h->Y);
    ZF_CUDA(cudaGetLastError());
    zf::zf_count_launch();
  }
  if (h->active) {
    h->phase = MP_GRAD;
    *next = 0;
    return ZF_OK;
  }
  bool all_known = true;
  for (int k = 0; k < h->n_runs; ++k)
    if (!h->run[k].F_known && !h->run[k].result_is_prev) all_known = false;
  if (all_known) {
    h->phase = MP_DONE;
    *next = 2;
  } else {
    h->phase = MP_FINAL;   // one more residual pass for res.fun = F(x) of every run
    *next = 1;
  }
  return ZF_OK;
}

// 2-D fp64 tensor map over a row-major array: boxes of 16 doubles (128 B, SWIZZLE_128B) x box_rows
// rows; elements outside [0, inner) x [0, outer) read as 0.  cuTensorMapEncodeTiled is taken from
// the driver through the runtime (no link-time dependency on libcuda).
int make_map(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch,
             int box_rows) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                               const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    ZF_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess)
      return zf::zf_fail(ZF_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  const cuuint64_t dims[2] = {inner, outer};
  const cuuint64_t strides[1] = {pitch * sizeof(double)};
  const cuuint32_t box[2] = {(cuuint32_t)zf::multi::BOXC, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void*>(base), dims,
                            strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return zf::zf_fail(ZF_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return ZF_OK;
}

int check_options(const zf_options* o) {
  if (!o) return zf::zf_fail(ZF_ERR_INVALID, "options is NULL");
  if (!(o->lr > 0.0)) return zf::zf_fail(ZF_ERR_INVALID, "lr must be > 0");
  if (o->max_iter < 1) return zf::zf_fail(ZF_ERR_INVALID, "max_iter must be >= 1");
  if (o->max_backtrack_iter < 1) return zf::zf_fail(ZF_ERR_INVALID, "max_backtrack_iter must be >= 1");
  if (!(o->decay_rate > 0.0 && o->decay_rate <= 1.0))
    return zf::zf_fail(ZF_ERR_INVALID, "decay_rate must be in (0, 1]");
  if (o->trace_capacity < 0) return zf::zf_fail(ZF_ERR_INVALID, "trace_capacity must be >= 0");
  return ZF_OK;
}

int begin_impl(zf_lasso_multi* h, const zf_options* opt, const double* d_x0, int x0_is_batched,
               const double* h_ab, double* h_allerrs, double* h_allfuns) {
  if (!h || !d_x0) return zf::zf_fail(ZF_ERR_INVALID, "NULL argument");
  int rc = check_options(opt);
  if (rc != ZF_OK) return rc;
  h->opt = *opt;
  h->h_allerrs = h_allerrs;
  h->h_allfuns = h_allfuns;
  h->need_F = (opt->decay_rate != 1.0) || (opt->trace_capacity > 0 && h_allfuns != nullptr);
  h->active = (h->n_runs == 32) ? 0xffffffffu : ((1u << h->n_runs) - 1u);
  h->trial = h->accepted = 0;
  for (int k = 0; k < h->n_runs; ++k) {
    RunState& r = h->run[k];
    r = RunState();
    r.lr = opt->lr;
    r.a = h_ab ? h_ab[2 * k] : opt->nesterov_a;
    r.b = h_ab ? h_ab[2 * k + 1] : opt->nesterov_b;
    r.err = INFINITY;
    const double* src = d_x0 + (x0_is_batched ? (long long)k * h->n_cols : 0);
    const size_t nb = sizeof(double) * (size_t)h->n_cols;
    const long long off = (long long)k * h->pitch_c;
    ZF_CUDA(cudaMemcpyAsync(h->Xp + off, src, nb, cudaMemcpyDeviceToDevice, h->st));
    ZF_CUDA(cudaMemcpyAsync(h->Xn + off, src, nb, cudaMemcpyDeviceToDevice, h->st));
    ZF_CUDA(cudaMemcpyAsync(h->Y + off, src, nb, cudaMemcpyDeviceToDevice, h->st));
  }
  // F(x0): residual norms (all-reduced by the caller if the rows are sharded)
  rc = launch_residual(h, 0);
  if (rc != ZF_OK) return rc;
  rc = launch_collect(h, false);
  if (rc != ZF_OK) return rc;
  h->phase = MP_INIT;
  return ZF_OK;
}

}  // namespace

extern "C" int zf_lasso_multi_create(zf_lasso_multi** out, const double* d_A, int64_t n_rows,
                                     int64_t n_cols, const double* d_b, int32_t b_is_batched,
                                     int32_t n_runs, double scale, double l1,
                                     void* cuda_stream) {
  if (!out || !d_A || !d_b) return zf::zf_fail(ZF_ERR_INVALID, "NULL argument");
  if (n_rows < 1 || n_cols < 1) return zf::zf_fail(ZF_ERR_INVALID, "n_rows and n_cols must be >= 1");
  if (n_runs < 1 || n_runs > zf::multi::MAX_RUNS)
    return zf::zf_fail(ZF_ERR_INVALID, "n_runs must be in 1..%d", zf::multi::MAX_RUNS);
  if ((n_cols & 1) || (reinterpret_cast<uintptr_t>(d_A) & 15u))
    return zf::zf_fail(ZF_ERR_UNSUPPORTED,
                       "the shared-A path streams rows with 16-byte bulk copies: n_cols must be "
                       "even and A 16-byte aligned");
  int rc = zf::zf_require_device();
  if (rc != ZF_OK) return rc;
  zf_lasso_multi* h = new (std::nothrow) zf_lasso_multi();
  if (!h) return zf::zf_fail(ZF_ERR_INVALID, "out of host memory");
  h->A = d_A;
  h->n_rows = n_rows;
  h->n_cols = n_cols;
  h->n_runs = n_runs;
  h->nt = (n_runs + 7) / 8;
  h->kp = 8 * h->nt;
  h->scale = scale;
  h->l1 = l1;
  h->st = (cudaStream_t)cuda_stream;
  h->b_batched = b_is_batched != 0;
  h->pitch_c = n_cols;                       // even
  h->pitch_r = (n_rows + 1) & ~1LL;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, dev);
  h->grid = h->n_sm;
  const long long n_slabs = (n_cols + zf::multi::TILE - 1) / zf::multi::TILE;
  const long long n_rowtiles = (n_rows + zf::multi::TILE - 1) / zf::multi::TILE;
  long long splits = (16LL * h->grid + n_slabs - 1) / n_slabs;     // >= 16 items per CTA
  if (splits > n_rowtiles) splits = n_rowtiles;
  if (splits < 1) splits = 1;
  h->rows_per_split = ((n_rowtiles + splits - 1) / splits) * zf::multi::TILE;
  h->n_splits = (int)((n_rows + h->rows_per_split - 1) / h->rows_per_split);
  auto fail = [&](int code) {
    zf_lasso_multi_destroy(h);
    return code;
  };
#define ZF_TRY(call)                                                      \
  do {                                                                    \
    cudaError_t _e = (call);                                              \
    if (_e != cudaSuccess) return fail(zf::zf_fail_cuda(_e, #call));      \
  } while (0)
  const size_t vec_b = sizeof(double) * (size_t)h->kp * (size_t)h->pitch_c;
  const size_t r_b = sizeof(double) * (size_t)h->kp * (size_t)h->pitch_r;
  ZF_TRY(cudaMalloc(&h->vecs, 4 * vec_b));
  ZF_TRY(cudaMemsetAsync(h->vecs, 0, 4 * vec_b, h->st));
  h->Xp = h->vecs;
  h->Xn = h->Xp + (size_t)h->kp * h->pitch_c;
  h->Y = h->Xn + (size_t)h->kp * h->pitch_c;
  h->G = h->Y + (size_t)h->kp * h->pitch_c;
  const size_t b_b = h->b_batched ? r_b : sizeof(double) * (size_t)h->pitch_r;
  ZF_TRY(cudaMalloc(&h->bcopy, b_b));
  ZF_TRY(cudaMemsetAsync(h->bcopy, 0, b_b, h->st));
  if (h->b_batched)
    ZF_TRY(cudaMemcpy2DAsync(h->bcopy, sizeof(double) * h->pitch_r, d_b, sizeof(double) * n_rows,
                             sizeof(double) * n_rows, n_runs, cudaMemcpyDeviceToDevice, h->st));
  else
    ZF_TRY(cudaMemcpyAsync(h->bcopy, d_b, sizeof(double) * n_rows, cudaMemcpyDeviceToDevice, h->st));
  ZF_TRY(cudaMalloc(&h->R, r_b));
  ZF_TRY(cudaMemsetAsync(h->R, 0, r_b, h->st));
  ZF_TRY(cudaMalloc(&h->gpart, vec_b * (size_t)h->n_splits));
  ZF_TRY(cudaMalloc(&h->sq_part, sizeof(double) * (size_t)h->grid * h->kp));
  ZF_TRY(cudaMalloc(&h->partial, vec_b + sizeof(double) * h->kp));
  ZF_TRY(cudaMemsetAsync(h->partial, 0, vec_b + sizeof(double) * h->kp, h->st));
  ZF_TRY(cudaMalloc(&h->block_sums, sizeof(StepSums) * (size_t)h->kp * zf::multi::VEC_BLOCKS));
  ZF_TRY(cudaMalloc(&h->d_sums, sizeof(StepSums) * h->kp));
  ZF_TRY(cudaMemsetAsync(h->d_sums, 0, sizeof(StepSums) * h->kp, h->st));
  ZF_TRY(cudaMalloc(&h->counter, sizeof(unsigned int) * h->kp));
  ZF_TRY(cudaMemsetAsync(h->counter, 0, sizeof(unsigned int) * h->kp, h->st));
  ZF_TRY(cudaMallocHost(&h->h_pin, sizeof(double) * 5 * h->kp));
  ZF_TRY(cudaStreamSynchronize(h->st));
#undef ZF_TRY
  rc = make_map(&h->tmA, d_A, (uint64_t)n_cols, (uint64_t)n_rows, (uint64_t)n_cols, zf::multi::TILE);
  if (rc == ZF_OK) rc = make_map(&h->tmY, h->Y, (uint64_t)n_cols, (uint64_t)h->kp, (uint64_t)h->pitch_c, h->kp);
  if (rc == ZF_OK) rc = make_map(&h->tmXn, h->Xn, (uint64_t)n_cols, (uint64_t)h->kp, (uint64_t)h->pitch_c, h->kp);
  if (rc == ZF_OK) rc = make_map(&h->tmR, h->R, (uint64_t)n_rows, (uint64_t)h->kp, (uint64_t)h->pitch_r, h->kp);
  if (rc != ZF_OK) return fail(rc);
  *out = h;
  return ZF_OK;
}

extern "C" void zf_lasso_multi_destroy(zf_lasso_multi* h) {
  if (!h) return;
  cudaFree(h->d_opts);
  cudaFree(h->d_state);
  cudaFree(h->d_allerrs);
  cudaFree(h->d_allfuns);
  if (h->h_state) cudaFreeHost(h->h_state);
  for (int k = 0; k < 2; ++k)
    if (h->ev_poll[k]) cudaEventDestroy(h->ev_poll[k]);
  cudaFree(h->vecs);
  cudaFree(h->bcopy);
  cudaFree(h->R);
  cudaFree(h->gpart);
  cudaFree(h->sq_part);
  cudaFree(h->partial);
  cudaFree(h->block_sums);
  cudaFree(h->d_sums);
  cudaFree(h->counter);
  if (h->h_pin) cudaFreeHost(h->h_pin);
  delete h;
}

extern "C" int zf_lasso_multi_begin(zf_lasso_multi* h, const zf_options* opt, const double* d_x0,
                                    int32_t x0_is_batched, const double* h_ab) {
  return begin_impl(h, opt, d_x0, x0_is_batched, h_ab, nullptr, nullptr);
}

extern "C" int zf_lasso_multi_grad(zf_lasso_multi* h, int which) {
  if (!h) return zf::zf_fail(ZF_ERR_INVALID, "NULL handle");
  if (which == 0) {
    if (h->phase != MP_GRAD) return zf::zf_fail(ZF_ERR_INVALID, "zf_lasso_multi_grad(0) out of order");
    return gradient_pass(h);
  }
  if (which == 1) {
    if (h->phase != MP_FNEW && h->phase != MP_FINAL)
      return zf::zf_fail(ZF_ERR_INVALID, "zf_lasso_multi_grad(1) out of order");
    int rc = launch_residual(h, 1);
    if (rc != ZF_OK) return rc;
    return launch_collect(h, false);
  }
  return zf::zf_fail(ZF_ERR_INVALID, "which must be 0 or 1");
}

/* the stream every later call enqueues on (callers whose current stream changes hand it over) */
extern "C" int zf_lasso_multi_set_stream(zf_lasso_multi* h, void* cuda_stream) {
  if (!h) return zf::zf_fail(ZF_ERR_INVALID, "NULL handle");
  if (h->phase != MP_IDLE && h->phase != MP_DONE)
    return zf::zf_fail(ZF_ERR_INVALID, "zf_lasso_multi_set_stream during a solve");
  h->st = (cudaStream_t)cuda_stream;
  return ZF_OK;
}

extern "C" double* zf_lasso_multi_partial(zf_lasso_multi* h, int64_t* n_values) {
  if (!h) return nullptr;
  if (n_values) *n_values = (int64_t)h->kp * h->pitch_c + h->kp;
  return h->partial;
}

extern "C" int zf_lasso_multi_step(zf_lasso_multi* h, int32_t* h_next) {
  if (!h || !h_next) return zf::zf_fail(ZF_ERR_INVALID, "NULL argument");
  int next = 2;
  int rc = ZF_OK;
  const int cap = h->opt.trace_capacity;
  switch (h->phase) {
    case MP_INIT: {
      // F(x0) = f(x0) + g(x0)   (proximal_gradient.py:466, 279)
      rc = run_prox(h, h->active, false, true);
      if (rc != ZF_OK) return rc;
      rc = fetch(h, true, h->active);
      if (rc != ZF_OK) return rc;
      for (int k = 0; k < h->n_runs; ++k) {
        RunState& r = h->run[k];
        r.F_prev = f_from_ss(h, ss_of(h, k)) + h->l1 * r.sums.abs1;
        r.F_x = r.F_prev;
        if (cap > 0 && h->h_allfuns) h->h_allfuns[(long long)k * (cap + 1)] = r.F_prev;
        r.nit = 1;
      }
      h->phase = MP_GRAD;
      next = 0;
      break;
    }
    case MP_GRAD: {
      h->trial = h->active;
      rc = run_prox(h, h->trial, true, false);
      if (rc != ZF_OK) return rc;
      rc = fetch(h, true, h->trial);
      if (rc != ZF_OK) return rc;
      for (int k = 0; k < h->n_runs; ++k) {
        if (!((h->trial >> k) & 1u)) continue;
        RunState& r = h->run[k];
        r.bt = 0;
        r.f_y = f_from_ss(h, ss_of(h, k));
        r.sub_fun = subproblem_fun(h, r);
        r.F_known = false;
      }
      if (h->need_F) {
        h->phase = MP_FNEW;
        next = 1;
      } else {
        h->accepted = h->trial;
        h->trial = 0;
        rc = advance(h, &next);
      }
      break;
    }
    case MP_FNEW: {
      rc = fetch(h, false, 0);
      if (rc != ZF_OK) return rc;
      for (int k = 0; k < h->n_runs; ++k) {
        if (!((h->trial >> k) & 1u)) continue;
        RunState& r = h->run[k];
        const double f_x = f_from_ss(h, ss_of(h, k));
        r.F_x = f_x + h->l1 * r.sums.abs1;
        r.F_known = true;
        bool ok;
        if (h->opt.decay_rate == 1.0) ok = true;                      // proximal_gradient.py:298
        else if (h->opt.deprecated) ok = (f_x - r.f_y <= r.sub_fun + h->opt.tol_internal);
        else ok = (r.F_x - r.F_prev <= r.sub_fun + h->opt.tol_internal);
        if (ok) {
          h->trial &= ~(1u << k);
          h->accepted |= 1u << k;
          continue;
        }
        r.lr *= h->opt.decay_rate;
        r.bt += 1;
        if (r.bt >= h->opt.max_backtrack_iter) {
          // RuntimeError("Backtracking failed ...") -> x = x_prev, nit - 1 (493-509)
          r.result_is_prev = true;
          r.F_x = r.F_prev;
          r.nit -= 1;
          r.status = -1;
          h->trial &= ~(1u << k);
          h->active &= ~(1u << k);
        }
      }
      if (h->trial) {
        // same gradients, smaller steps: redo the prox of the rejected runs from the stored g
        rc = run_prox(h, h->trial, false, false);
        if (rc != ZF_OK) return rc;
        rc = fetch(h, true, h->trial);
        if (rc != ZF_OK) return rc;
        for (int k = 0; k < h->n_runs; ++k)
          if ((h->trial >> k) & 1u) h->run[k].sub_fun = subproblem_fun(h, h->run[k]);
        h->phase = MP_FNEW;
        next = 1;
      } else {
        rc = advance(h, &next);
      }
      break;
    }
    case MP_FINAL: {
      rc = fetch(h, false, 0);
      if (rc != ZF_OK) return rc;
      for (int k = 0; k < h->n_runs; ++k) {
        RunState& r = h->run[k];
        if (r.F_known || r.result_is_prev) continue;
        r.F_x = f_from_ss(h, ss_of(h, k)) + h->l1 * r.sums.abs1;
        r.F_known = true;
      }
      h->phase = MP_DONE;
      next = 2;
      break;
    }
    case MP_DONE:
      next = 2;
      break;
    default:
      return zf::zf_fail(ZF_ERR_INVALID, "zf_lasso_multi_step called before zf_lasso_multi_begin");
  }
  if (rc != ZF_OK) return rc;
  *h_next = next;
  return ZF_OK;
}

extern "C" int zf_lasso_multi_finish(zf_lasso_multi* h, double* d_x, double* h_fun, int64_t* h_nit,
                                     int32_t* h_status, double* h_lr, double* h_err) {
  if (!h) return zf::zf_fail(ZF_ERR_INVALID, "NULL handle");
  if (h->phase != MP_DONE)
    return zf::zf_fail(ZF_ERR_INVALID, "zf_lasso_multi_finish before the solve ended");
  for (int k = 0; k < h->n_runs; ++k) {
    const RunState& r = h->run[k];
    if (d_x) {
      // a run that ended normally keeps its result in the candidate buffer; after a failed
      // line search the result is the previous iterate
      const double* prev = h->Xp + (long long)k * h->pitch_c;
      const double* cand = h->Xn + (long long)k * h->pitch_c;
      ZF_CUDA(cudaMemcpyAsync(d_x + (long long)k * h->n_cols, r.result_is_prev ? prev : cand,
                              sizeof(double) * (size_t)h->n_cols, cudaMemcpyDeviceToDevice, h->st));
    }
    if (h_fun) h_fun[k] = r.F_x;
    if (h_nit) h_nit[k] = r.nit;
    if (h_status) h_status[k] = r.status;
    if (h_lr) h_lr[k] = r.lr;
    if (h_err) h_err[k] = r.err;
  }
  if (d_x) ZF_CUDA(cudaStreamSynchronize(h->st));
  h->phase = MP_IDLE;
  return ZF_OK;
}

// ---- device-decided rounds (single GPU): nothing in the loop waits for the GPU except the poll
namespace {

int multi_dev_alloc(zf_lasso_multi* h) {
  if (h->d_state) return ZF_OK;
  ZF_CUDA(cudaMalloc((void**)&h->d_opts, sizeof(zf::multi::MultiDevOpts)));
  ZF_CUDA(cudaMalloc((void**)&h->d_state, sizeof(zf::multi::MultiDevState)));
  ZF_CUDA(cudaMemset(h->d_state, 0, sizeof(zf::multi::MultiDevState)));
  ZF_CUDA(cudaMallocHost((void**)&h->h_state, 2 * sizeof(zf::multi::MultiDevState)));
  for (int k = 0; k < 2; ++k)
    ZF_CUDA(cudaEventCreateWithFlags(&h->ev_poll[k], cudaEventDisableTiming));
  return ZF_OK;
}

// Stages of a device-decided solve (the numbering of zf_lasso_dev_stage in zf_lasso.cu).  A trial
// ("slot") is GRAD, PROX and -- with a line search or an F trace -- FEVAL, DECIDE.  A row-sharded
// caller all-reduces `partial` after GRAD and its residual-norm tail after begin, FEVAL and
// FEVAL_FINAL; every kernel that reads those values is skipped or reads its stored copy when the
// pass before the exchange was skipped, so the exchange is unconditional.
enum { MS_INIT = 0, MS_GRAD = 1, MS_PROX = 2, MS_FEVAL = 3, MS_DECIDE = 4, MS_FINAL = 5,
       MS_FEVAL_FINAL = 6 };

int multi_dev_stage(zf_lasso_multi* h, int stage) {
  using namespace zf::multi;
  const double* ss = h->partial + (long long)h->kp * h->pitch_c;
  dim3 vgrid(VEC_BLOCKS, (unsigned)h->n_runs);
  int rc = ZF_OK;
  switch (stage) {
    case MS_INIT:      // |x0|_1 of every run, then F(x0) from the (reduced) residual norms
      rc = run_prox(h, h->active, false, true);
      if (rc != ZF_OK) return rc;
      multi_dev_init_kernel<<<1, 32, 0, h->st>>>(h->d_opts, h->d_state, h->d_sums, ss);
      ZF_CUDA(cudaGetLastError());
      zf::zf_count_launch();
      return ZF_OK;
    case MS_GRAD:      // both DGEMM passes + collect (or nothing while a retry is pending)
      h->skip = &h->d_state->skip_grad;
      rc = gradient_pass(h);
      h->skip = nullptr;
      return rc;
    case MS_PROX:
      multi_dev_prox_kernel<<<vgrid, VEC_THREADS, 0, h->st>>>(h->d_opts, h->d_state, h->Y, h->Xn,
                                                              h->partial, h->G, h->block_sums,
                                                              h->counter, h->d_sums);
      ZF_CUDA(cudaGetLastError());
      zf::zf_count_launch();
      multi_dev_after_prox_kernel<<<1, 32, 0, h->st>>>(h->d_opts, h->d_state, h->d_sums, ss);
      ZF_CUDA(cudaGetLastError());
      zf::zf_count_launch();
      if (h->dev_need_F) return ZF_OK;
      break;           // fixed step, no F trace: the trial is accepted, momentum follows
    case MS_FEVAL:
      h->skip = &h->d_state->done;
      rc = launch_residual(h, 1);
      h->skip = nullptr;
      if (rc != ZF_OK) return rc;
      return launch_collect(h, false);
    case MS_DECIDE:
      multi_dev_decide_kernel<<<1, 32, 0, h->st>>>(h->d_opts, h->d_state, ss);
      ZF_CUDA(cudaGetLastError());
      zf::zf_count_launch();
      break;
    case MS_FEVAL_FINAL:   // res.fun = F(x) of runs that never evaluated it: one residual pass
      rc = launch_residual(h, 1);
      if (rc != ZF_OK) return rc;
      return launch_collect(h, false);
    case MS_FINAL:
      multi_dev_final_kernel<<<1, 32, 0, h->st>>>(h->d_opts, h->d_state, ss);
      ZF_CUDA(cudaGetLastError());
      zf::zf_count_launch();
      return ZF_OK;
    default:
      return zf::zf_fail(ZF_ERR_INVALID, "unknown stage %d", stage);
  }
  multi_dev_momentum_kernel<<<vgrid, VEC_THREADS, 0, h->st>>>(h->d_opts, h->d_state, h->Xp, h->Xn, h->Y);
  ZF_CUDA(cudaGetLastError());
  zf::zf_count_launch();
  return ZF_OK;
}

// one trial of the runs on trial (single GPU: no exchange between the stages)
int multi_dev_slot(zf_lasso_multi* h) {
  int rc = multi_dev_stage(h, MS_GRAD);
  if (rc == ZF_OK) rc = multi_dev_stage(h, MS_PROX);
  if (rc == ZF_OK && h->dev_need_F) {
    rc = multi_dev_stage(h, MS_FEVAL);
    if (rc == ZF_OK) rc = multi_dev_stage(h, MS_DECIDE);
  }
  return rc;
}

// options and starts to the device, F(x0)'s residual pass enqueued (its norms are in `partial`)
int multi_dev_begin(zf_lasso_multi* h, const zf_options* opt, const double* d_x0,
                    int32_t x0_is_batched, const double* h_ab, bool want_errs, bool want_funs) {
  using namespace zf::multi;
  if (!h || !d_x0) return zf::zf_fail(ZF_ERR_INVALID, "NULL argument");
  int rc = check_options(opt);
  if (rc != ZF_OK) return rc;
  rc = multi_dev_alloc(h);
  if (rc != ZF_OK) return rc;
  h->opt = *opt;
  const int cap = (want_errs || want_funs) ? opt->trace_capacity : 0;
  const size_t need = (size_t)h->n_runs * ((size_t)cap + 1);
  if (cap > 0 && need > h->dev_trace) {
    cudaFree(h->d_allerrs);
    cudaFree(h->d_allfuns);
    h->d_allerrs = h->d_allfuns = nullptr;
    h->dev_trace = 0;
    ZF_CUDA(cudaMalloc((void**)&h->d_allerrs, sizeof(double) * need));
    ZF_CUDA(cudaMalloc((void**)&h->d_allfuns, sizeof(double) * need));
    h->dev_trace = need;
  }
  MultiDevOpts o{};
  o.lr0 = opt->lr; o.tol = opt->tol; o.tol_internal = opt->tol_internal; o.decay = opt->decay_rate;
  o.scale = h->scale; o.l1 = h->l1; o.max_iter = opt->max_iter; o.pitch_c = h->pitch_c;
  o.n = h->n_cols; o.max_bt = opt->max_backtrack_iter; o.nesterov = opt->nesterov;
  o.deprecated = opt->deprecated; o.cap = cap; o.n_runs = h->n_runs; o.kp = h->kp;
  o.need_F = (opt->decay_rate != 1.0) || (cap > 0 && want_funs);
  o.allerrs = (cap > 0 && want_errs) ? h->d_allerrs : nullptr;
  o.allfuns = (cap > 0 && want_funs) ? h->d_allfuns : nullptr;
  for (int k = 0; k < h->n_runs; ++k) {
    o.na[k] = h_ab ? h_ab[2 * k] : opt->nesterov_a;
    o.nb[k] = h_ab ? h_ab[2 * k + 1] : opt->nesterov_b;
  }
  h->dev_need_F = o.need_F;
  h->dev_cap = cap;
  ZF_CUDA(cudaMemcpyAsync(h->d_opts, &o, sizeof(o), cudaMemcpyHostToDevice, h->st));
  if (cap > 0) {
    if (o.allerrs) ZF_CUDA(cudaMemsetAsync(h->d_allerrs, 0, sizeof(double) * (size_t)h->n_runs * cap, h->st));
    if (o.allfuns) ZF_CUDA(cudaMemsetAsync(h->d_allfuns, 0, sizeof(double) * need, h->st));
  }
  for (int k = 0; k < h->n_runs; ++k) {
    const double* src = d_x0 + (x0_is_batched ? (long long)k * h->n_cols : 0);
    const size_t nb = sizeof(double) * (size_t)h->n_cols;
    const long long off = (long long)k * h->pitch_c;
    ZF_CUDA(cudaMemcpyAsync(h->Xp + off, src, nb, cudaMemcpyDeviceToDevice, h->st));
    ZF_CUDA(cudaMemcpyAsync(h->Xn + off, src, nb, cudaMemcpyDeviceToDevice, h->st));
    ZF_CUDA(cudaMemcpyAsync(h->Y + off, src, nb, cudaMemcpyDeviceToDevice, h->st));
  }
  h->active = (h->n_runs == 32) ? 0xffffffffu : ((1u << h->n_runs) - 1u);
  for (int k = 0; k < h->n_runs; ++k) h->run[k].lr = opt->lr;     // run_prox reads the host copy
  // F(x0): residual norms of every run (all-reduced by the caller if the rows are sharded)
  rc = launch_residual(h, 0);
  if (rc == ZF_OK) rc = launch_collect(h, false);
  if (rc != ZF_OK) return rc;
  h->phase = MP_INIT;
  return ZF_OK;
}

int multi_dev_snapshot(zf_lasso_multi* h, int slot) {
  ZF_CUDA(cudaMemcpyAsync(&h->h_state[slot], h->d_state, sizeof(zf::multi::MultiDevState),
                          cudaMemcpyDeviceToHost, h->st));
  ZF_CUDA(cudaEventRecord(h->ev_poll[slot], h->st));
  return ZF_OK;
}

int multi_dev_finish(zf_lasso_multi* h, double* d_x, double* h_fun, int64_t* h_nit,
                     int32_t* h_status, double* h_lr, double* h_err, double* h_allerrs,
                     double* h_allfuns) {
  using namespace zf::multi;
  const int cap = h->dev_cap;
  const size_t need = (size_t)h->n_runs * ((size_t)cap + 1);
  ZF_CUDA(cudaMemcpyAsync(&h->h_state[0], h->d_state, sizeof(MultiDevState), cudaMemcpyDeviceToHost,
                          h->st));
  ZF_CUDA(cudaStreamSynchronize(h->st));
  const MultiDevState& S = h->h_state[0];
  for (int k = 0; k < h->n_runs; ++k) {
    const MultiDevRun& r = S.run[k];
    if (d_x) {
      const double* prev = h->Xp + (long long)k * h->pitch_c;
      const double* cand = h->Xn + (long long)k * h->pitch_c;
      ZF_CUDA(cudaMemcpyAsync(d_x + (long long)k * h->n_cols, r.result_is_prev ? prev : cand,
                              sizeof(double) * (size_t)h->n_cols, cudaMemcpyDeviceToDevice, h->st));
    }
    if (h_fun) h_fun[k] = r.F_x;
    if (h_nit) h_nit[k] = r.nit;
    if (h_status) h_status[k] = r.status;
    if (h_lr) h_lr[k] = r.lr;
    if (h_err) h_err[k] = r.err;
  }
  if (cap > 0 && h_allerrs)
    ZF_CUDA(cudaMemcpyAsync(h_allerrs, h->d_allerrs, sizeof(double) * (size_t)h->n_runs * cap,
                            cudaMemcpyDeviceToHost, h->st));
  if (cap > 0 && h_allfuns)
    ZF_CUDA(cudaMemcpyAsync(h_allfuns, h->d_allfuns, sizeof(double) * need, cudaMemcpyDeviceToHost, h->st));
  ZF_CUDA(cudaStreamSynchronize(h->st));
  h->phase = MP_IDLE;
  return ZF_OK;
}

int multi_solve_dev(zf_lasso_multi* h, const zf_options* opt, const double* d_x0,
                    int32_t x0_is_batched, const double* h_ab, double* d_x, double* h_fun,
                    int64_t* h_nit, int32_t* h_status, double* h_lr, double* h_err,
                    double* h_allerrs, double* h_allfuns) {
  int rc = multi_dev_begin(h, opt, d_x0, x0_is_batched, h_ab, h_allerrs != nullptr,
                           h_allfuns != nullptr);
  if (rc == ZF_OK) rc = multi_dev_stage(h, MS_INIT);
  // chunks of slots, the poll one chunk behind
  const int chunk = 8;
  int cur = 0;
  for (int k = 0; k < chunk && rc == ZF_OK; ++k) rc = multi_dev_slot(h);
  if (rc == ZF_OK) rc = multi_dev_snapshot(h, cur);
  while (rc == ZF_OK) {
    for (int k = 0; k < chunk && rc == ZF_OK; ++k) rc = multi_dev_slot(h);
    if (rc == ZF_OK) rc = multi_dev_snapshot(h, 1 - cur);
    if (rc != ZF_OK) break;
    ZF_CUDA(cudaEventSynchronize(h->ev_poll[cur]));
    if (h->h_state[cur].done) break;
    cur = 1 - cur;
  }
  if (rc == ZF_OK && !h->dev_need_F) rc = multi_dev_stage(h, MS_FEVAL_FINAL);
  if (rc == ZF_OK) rc = multi_dev_stage(h, MS_FINAL);
  if (rc != ZF_OK) {
    if (h) h->phase = MP_IDLE;
    return rc;
  }
  return multi_dev_finish(h, d_x, h_fun, h_nit, h_status, h_lr, h_err, h_allerrs, h_allfuns);
}

}  // namespace

extern "C" int zf_lasso_multi_solve(zf_lasso_multi* h, const zf_options* opt, const double* d_x0,
                                    int32_t x0_is_batched, const double* h_ab, double* d_x,
                                    double* h_fun, int64_t* h_nit, int32_t* h_status,
                                    double* h_lr, double* h_err, double* h_allerrs,
                                    double* h_allfuns) {
  // default: the device-decided rounds; ZF_LASSO_HOSTLOOP=1 keeps the round-1 loop (decisions on
  // the host, one D2H copy + stream sync per trial) for A/B measurements
  static const bool hostloop = getenv("ZF_LASSO_HOSTLOOP") && getenv("ZF_LASSO_HOSTLOOP")[0] == '1';
  if (!hostloop)
    return multi_solve_dev(h, opt, d_x0, x0_is_batched, h_ab, d_x, h_fun, h_nit, h_status, h_lr,
                           h_err, h_allerrs, h_allfuns);
  int rc = begin_impl(h, opt, d_x0, x0_is_batched, h_ab, h_allerrs, h_allfuns);
  if (rc != ZF_OK) return rc;
  int32_t next = 0;
  rc = zf_lasso_multi_step(h, &next);
  while (rc == ZF_OK && next != 2) {
    rc = zf_lasso_multi_grad(h, next);
    if (rc != ZF_OK) break;
    rc = zf_lasso_multi_step(h, &next);
  }
  if (rc != ZF_OK) {
    h->phase = MP_IDLE;
    return rc;
  }
  return zf_lasso_multi_finish(h, d_x, h_fun, h_nit, h_status, h_lr, h_err);
}

// ---- the device-decided solve in pieces, for a caller that owns an exchange between the stages
// (row-sharded runs: zfista_b200/distributed.py run_device_lasso drives these exactly as it drives
// zf_lasso_dev_*; same stage numbers)
extern "C" int zf_lasso_multi_dev_begin(zf_lasso_multi* h, const zf_options* opt,
                                        const double* d_x0, int32_t x0_is_batched,
                                        const double* h_ab) {
  return multi_dev_begin(h, opt, d_x0, x0_is_batched, h_ab, false, false);
}
extern "C" int zf_lasso_multi_dev_stage(zf_lasso_multi* h, int32_t stage) {
  if (!h) return zf::zf_fail(ZF_ERR_INVALID, "NULL handle");
  if (h->phase != MP_INIT) return zf::zf_fail(ZF_ERR_INVALID, "zf_lasso_multi_dev_stage outside a solve");
  return multi_dev_stage(h, stage);
}
extern "C" int zf_lasso_multi_dev_needs_feval(zf_lasso_multi* h) { return h ? h->dev_need_F : 0; }
extern "C" int zf_lasso_multi_dev_poll(zf_lasso_multi* h, int32_t slot, int32_t wait, int32_t* done) {
  if (!h || slot < 0 || slot > 1) return zf::zf_fail(ZF_ERR_INVALID, "bad argument");
  if (!wait) return multi_dev_snapshot(h, slot);
  ZF_CUDA(cudaEventSynchronize(h->ev_poll[slot]));
  if (done) *done = h->h_state[slot].done;
  return ZF_OK;
}
extern "C" int zf_lasso_multi_dev_finish(zf_lasso_multi* h, double* d_x, double* h_fun,
                                         int64_t* h_nit, int32_t* h_status, double* h_lr,
                                         double* h_err) {
  if (!h) return zf::zf_fail(ZF_ERR_INVALID, "NULL handle");
  if (h->phase != MP_INIT) return zf::zf_fail(ZF_ERR_INVALID, "zf_lasso_multi_dev_finish outside a solve");
  return multi_dev_finish(h, d_x, h_fun, h_nit, h_status, h_lr, h_err, nullptr, nullptr);
}
/* `partial` is [kp][pitch_c] gradients followed by kp residual norms */
extern "C" int zf_lasso_multi_layout(zf_lasso_multi* h, int64_t* kp, int64_t* pitch_c) {
  if (!h) return zf::zf_fail(ZF_ERR_INVALID, "NULL handle");
  if (kp) *kp = h->kp;
  if (pitch_c) *pitch_c = h->pitch_c;
  return ZF_OK;
}

// the closure entries evaluate at caller-supplied points: they go through the y_k array (the
// array the vector tensor map covers), so they are only legal between solves
static int load_points(zf_lasso_multi* h, const double* d_X) {
  if (h->phase != MP_IDLE && h->phase != MP_DONE)
    return zf::zf_fail(ZF_ERR_INVALID, "closure evaluation in the middle of a solve");
  ZF_CUDA(cudaMemcpyAsync(h->Y, d_X, sizeof(double) * (size_t)h->n_runs * (size_t)h->n_cols,
                          cudaMemcpyDeviceToDevice, h->st));
  return ZF_OK;
}

extern "C" int zf_lasso_multi_gradient_device(zf_lasso_multi* h, const double* d_X,
                                              double* d_grad, double* d_f) {
  if (!h || !d_X || !d_grad) return zf::zf_fail(ZF_ERR_INVALID, "NULL argument");
  int rc = load_points(h, d_X);
  if (rc != ZF_OK) return rc;
  rc = gradient_pass(h);
  if (rc != ZF_OK) return rc;
  dim3 grid(zf::multi::VEC_BLOCKS, (unsigned)h->n_runs);
  zf::multi::multi_scale_kernel<<<grid, zf::multi::VEC_THREADS, 0, h->st>>>(
      h->partial, h->kp, h->pitch_c, h->n_cols, 2.0 * h->scale, h->scale, d_grad, d_f);
  ZF_CUDA(cudaGetLastError());
  zf::zf_count_launch();
  return ZF_OK;
}

extern "C" int zf_lasso_multi_pass_device(zf_lasso_multi* h, const double* d_X, int which) {
  // one pass alone (bench / roofline): which = 0 residual DGEMM of X, 1 the A^T R DGEMM of the
  // residuals the last pass-0 left in the workspace
  if (!h) return zf::zf_fail(ZF_ERR_INVALID, "NULL handle");
  if (which == 0) {
    if (!d_X) return zf::zf_fail(ZF_ERR_INVALID, "NULL argument");
    const int rc = load_points(h, d_X);
    if (rc != ZF_OK) return rc;
    return launch_residual(h, 0);
  }
  return launch_atr(h);
}
