// Batched FISTA / ISTA: host-side dispatch and the C ABI of the batched entry points.  The
// kernels (the whole minimize_proximal_gradient loop, proximal_gradient.py:474-555, on device,
// one warp per starting point) are in zf_batched_kernels.cuh.
#include <cstdio>
#include <cstring>
#include <mutex>

#include "zf_batched_kernels.cuh"

namespace zf {

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
  if ((P.kind == ZF_SD || P.kind == ZF_ZDT1) && (!P.has_bounds || P.has_l1))
    return zf_fail(ZF_ERR_INVALID, "SD / ZDT1 are defined with bounds (1e-6, inf) and without "
                                   "l1 terms (problems.py:238-245, 365-366)");
  if (P.has_bounds && P.bounds_are_arrays && (!P.lower_v || !P.upper_v))
    return zf_fail(ZF_ERR_INVALID, "array bounds requested but lower_v/upper_v missing");
  return ZF_OK;
}

// launch_t<KIND, M, GF> is instantiated in the zf_batched_inst_*.cu files (one translation unit
// per group of problem classes, so that the build compiles them in parallel).  GF = which parts
// g has (ZF_G_L1 | ZF_G_BOX).
#define ZF_DECL(K, M_, G) extern template int launch_t<K, M_, G>(const LaunchArgs&);
#define ZF_DECL4(K, M_) ZF_DECL(K, M_, 0) ZF_DECL(K, M_, 1) ZF_DECL(K, M_, 2) ZF_DECL(K, M_, 3)
ZF_DECL4(ZF_JOS1, 2)
ZF_DECL(ZF_SD, 2, ZF_G_BOX)
ZF_DECL4(ZF_FDS, 3)
ZF_DECL(ZF_ZDT1, 2, ZF_G_BOX)
ZF_DECL4(ZF_TOI4, 2)
ZF_DECL4(ZF_TRIDIA, 3)
ZF_DECL4(ZF_LFR1, 1) ZF_DECL4(ZF_LFR1, 2) ZF_DECL4(ZF_LFR1, 3) ZF_DECL4(ZF_LFR1, 4)
ZF_DECL(ZF_LSQ_L1, 1, 0) ZF_DECL(ZF_LSQ_L1, 2, 0) ZF_DECL(ZF_LSQ_L1, 3, 0)
#undef ZF_DECL4
#undef ZF_DECL

template <int KIND, int M>
static int launch_g(const LaunchArgs& L) {
  switch ((L.P.has_l1 ? ZF_G_L1 : 0) | (L.P.has_bounds ? ZF_G_BOX : 0)) {
    case 0: return launch_t<KIND, M, 0>(L);
    case 1: return launch_t<KIND, M, 1>(L);
    case 2: return launch_t<KIND, M, 2>(L);
    default: return launch_t<KIND, M, 3>(L);
  }
}

int zf_launch(const LaunchArgs& L) {
  int rc = validate_problem(L.P);
  if (rc != ZF_OK) return rc;
  const int m = L.P.n_objectives;
  switch (L.P.kind) {
    case ZF_JOS1: return launch_g<ZF_JOS1, 2>(L);
    case ZF_SD: return launch_t<ZF_SD, 2, ZF_G_BOX>(L);   // SD / ZDT1: box (1e-6, inf), no l1
    case ZF_FDS: return launch_g<ZF_FDS, 3>(L);
    case ZF_ZDT1: return launch_t<ZF_ZDT1, 2, ZF_G_BOX>(L);
    case ZF_TOI4: return launch_g<ZF_TOI4, 2>(L);
    case ZF_TRIDIA: return launch_g<ZF_TRIDIA, 3>(L);
    case ZF_LFR1:
      if (m == 1) return launch_g<ZF_LFR1, 1>(L);
      if (m == 2) return launch_g<ZF_LFR1, 2>(L);
      if (m == 3) return launch_g<ZF_LFR1, 3>(L);
      return launch_g<ZF_LFR1, 4>(L);
    case ZF_LSQ_L1:
      if (m == 1) return launch_t<ZF_LSQ_L1, 1, 0>(L);
      if (m == 2) return launch_t<ZF_LSQ_L1, 2, 0>(L);
      return launch_t<ZF_LSQ_L1, 3, 0>(L);
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

// Device scratch of the *_host entry points: one grow-only block PER CALLING THREAD, carved per
// call (cudaMalloc / cudaFree cost ~0.3 ms each and cudaFree synchronises; a 1024-start solve is
// ~1 ms of kernel), with its own non-blocking stream: host calls from different threads neither
// serialise on a lock nor on the legacy default stream (round 1: one process-wide arena, a
// mutex, stream 0).
struct Arena {
  cudaStream_t stream = nullptr;
  cudaStream_t get_stream() {
    if (!stream && cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess) {
      cudaGetLastError();
      stream = nullptr;              // fall back to the default stream
    }
    return stream;
  }
  ~Arena() {                         // thread exit (errors at process teardown are harmless)
    if (base) cudaFree(base);
    if (h_stage) cudaFreeHost(h_stage);
    if (stream) cudaStreamDestroy(stream);
    cudaGetLastError();
  }
  char* base = nullptr;
  size_t cap = 0, off = 0;
  int device = -1;
  // pinned host staging for the per-start results: ONE device-to-host copy per call instead of
  // eight small ones into pageable memory (each of those is its own synchronous transfer)
  char* h_stage = nullptr;
  size_t h_cap = 0;
  char* stage(size_t need) {
    if (need > h_cap) {
      if (h_stage) cudaFreeHost(h_stage);
      h_stage = nullptr;
      h_cap = 0;
      const size_t want = need + (need >> 2) + (1u << 16);
      if (cudaMallocHost((void**)&h_stage, want) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
      }
      h_cap = want;
    }
    return h_stage;
  }
  int reset(size_t need) {
    int dev = 0;
    ZF_CUDA(cudaGetDevice(&dev));
    if (dev != device || need > cap) {
      if (base) {
        cudaSetDevice(device);
        cudaFree(base);
        if (stream && dev != device) { cudaStreamDestroy(stream); stream = nullptr; }
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
thread_local Arena g_arena;
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
  cudaStream_t st = nullptr;
  const size_t N = (size_t)n_starts, n = problem->n_features, m = problem->n_objectives;
  const size_t cap = (size_t)opt->trace_capacity;
  // trace sizes: dense (cap entries per start) or ragged (h_out->trace_offsets, host array)
  const int64_t* h_off = h_out->trace_offsets;
  if (h_off && (h_off[0] != 0 || h_off[N] < 0))
    return zf::zf_fail(ZF_ERR_INVALID, "trace_offsets must start at 0 and be non-decreasing");
  const bool tracing = h_off != nullptr || cap > 0;
  const size_t n_err = h_off ? (size_t)h_off[N] : N * cap;          // error entries
  const size_t n_fx = h_off ? (size_t)h_off[N] + N : N * (cap + 1);  // F / x entries
  {
    size_t total = DevProblem::bytes(*problem) + 2 * pad256(N * n * 8) + pad256(N * 16) +
                   pad256(N * m * 8) + 6 * pad256(N * 8) + pad256((N + 1) * 8);
    if (tracing) {
      if (h_out->allerrs) total += pad256(n_err * 8);
      if (h_out->allfuns) total += pad256(n_fx * m * 8);
      if (h_out->allvecs) total += pad256(n_fx * n * 8);
    }
    rc = g_arena.reset(total);
    if (rc != ZF_OK) return rc;
  }
  st = g_arena.get_stream();
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
  const size_t res_begin = g_arena.off;     // x .. err are carved back to back: one D2H range
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
  const size_t res_end = g_arena.off;
  DevBuf toff;
  if (tracing) {
    if (h_off) {
      toff.take((N + 1) * 8);
      ZF_CUDA(cudaMemcpyAsync(toff.p, h_off, (N + 1) * 8, cudaMemcpyHostToDevice, st));
      R.trace_offsets = toff.as<int64_t>();
    }
    if (h_out->allerrs) {
      allerrs.take(n_err * 8);
      ZF_CUDA(cudaMemsetAsync(allerrs.p, 0, n_err * 8, st));
      R.allerrs = allerrs.as<double>();
    }
    if (h_out->allfuns) {
      allfuns.take(n_fx * m * 8);
      ZF_CUDA(cudaMemsetAsync(allfuns.p, 0, n_fx * m * 8, st));
      R.allfuns = allfuns.as<double>();
    }
    if (h_out->allvecs) {
      allvecs.take(n_fx * n * 8);
      ZF_CUDA(cudaMemsetAsync(allvecs.p, 0, n_fx * n * 8, st));
      R.allvecs = allvecs.as<double>();
    }
  }
  rc = zf_solve_batched_device(&dp.P, opt, n_starts, x0.as<double>(),
                               h_ab ? ab.as<double>() : nullptr, &R, st);
  if (rc != ZF_OK) return rc;
  char* stage = g_arena.stage(res_end - res_begin);
  if (stage) {
    // one transfer of the whole result block into pinned memory, then host copies
    ZF_CUDA(cudaMemcpyAsync(stage, g_arena.base + res_begin, res_end - res_begin,
                            cudaMemcpyDeviceToHost, st));
  } else {
    ZF_CUDA(cudaMemcpyAsync(h_out->x, R.x, N * n * 8, cudaMemcpyDeviceToHost, st));
    ZF_CUDA(cudaMemcpyAsync(h_out->fun, R.fun, N * m * 8, cudaMemcpyDeviceToHost, st));
    ZF_CUDA(cudaMemcpyAsync(h_out->nit, R.nit, N * 8, cudaMemcpyDeviceToHost, st));
    ZF_CUDA(cudaMemcpyAsync(h_out->status, R.status, N * 4, cudaMemcpyDeviceToHost, st));
    if (R.lr) ZF_CUDA(cudaMemcpyAsync(h_out->lr, R.lr, N * 8, cudaMemcpyDeviceToHost, st));
    if (R.nfev) ZF_CUDA(cudaMemcpyAsync(h_out->nfev, R.nfev, N * 8, cudaMemcpyDeviceToHost, st));
    if (R.n_dual) ZF_CUDA(cudaMemcpyAsync(h_out->n_dual, R.n_dual, N * 8, cudaMemcpyDeviceToHost, st));
    if (R.err) ZF_CUDA(cudaMemcpyAsync(h_out->err, R.err, N * 8, cudaMemcpyDeviceToHost, st));
  }
  if (R.allerrs && n_err) ZF_CUDA(cudaMemcpyAsync(h_out->allerrs, R.allerrs, n_err * 8, cudaMemcpyDeviceToHost, st));
  if (R.allfuns) ZF_CUDA(cudaMemcpyAsync(h_out->allfuns, R.allfuns, n_fx * m * 8, cudaMemcpyDeviceToHost, st));
  if (R.allvecs) ZF_CUDA(cudaMemcpyAsync(h_out->allvecs, R.allvecs, n_fx * n * 8, cudaMemcpyDeviceToHost, st));
  ZF_CUDA(cudaStreamSynchronize(st));
  if (stage) {
    auto back = [&](void* dst, const void* dev, size_t bytes) {
      std::memcpy(dst, stage + (static_cast<const char*>(dev) - (g_arena.base + res_begin)), bytes);
    };
    back(h_out->x, R.x, N * n * 8);
    back(h_out->fun, R.fun, N * m * 8);
    back(h_out->nit, R.nit, N * 8);
    back(h_out->status, R.status, N * 4);
    if (R.lr) back(h_out->lr, R.lr, N * 8);
    if (R.nfev) back(h_out->nfev, R.nfev, N * 8);
    if (R.n_dual) back(h_out->n_dual, R.n_dual, N * 8);
    if (R.err) back(h_out->err, R.err, N * 8);
  }
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
  cudaStream_t st = nullptr;
  const size_t N = (size_t)n, nf = problem->n_features, m = problem->n_objectives;
  rc = g_arena.reset(DevProblem::bytes(*problem) + 3 * pad256(N * nf * 8) + 3 * pad256(N * 8) +
                     pad256(N * m * 8));
  if (rc != ZF_OK) return rc;
  st = g_arena.get_stream();
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
  cudaStream_t st = nullptr;
  const size_t N = (size_t)n, nf = problem->n_features, m = problem->n_objectives;
  rc = g_arena.reset(DevProblem::bytes(*problem) + 2 * pad256(N * nf * 8) + 3 * pad256(N * m * 8) +
                     pad256(N * m * nf * 8));
  if (rc != ZF_OK) return rc;
  st = g_arena.get_stream();
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
