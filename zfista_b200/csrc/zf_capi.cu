// Library-wide pieces of the C ABI: versioning, error reporting, defaults.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "zf_host.h"

namespace zf {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

int zf_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int zf_fail_cuda(cudaError_t e, const char* what) {
  snprintf(g_err, sizeof(g_err), "CUDA error in %s: %s (%s)", what, cudaGetErrorName(e),
           cudaGetErrorString(e));
  return ZF_ERR_CUDA;
}

int zf_require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return zf_fail_cuda(e, "cudaGetDeviceCount (zfista_b200 has no CPU fallback)");
  }
  if (n < 1) return zf_fail(ZF_ERR_CUDA, "no CUDA device visible (zfista_b200 has no CPU fallback)");
  return ZF_OK;
}

void zf_count_launch(int64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace zf

extern "C" int zf_abi_version(void) { return ZF_ABI_VERSION; }

extern "C" const char* zf_last_error(void) { return zf::g_err; }

extern "C" int zf_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" int64_t zf_launch_count(void) { return zf::g_launches.load(); }

// defaults of minimize_proximal_gradient (proximal_gradient.py:311-331)
extern "C" void zf_default_options(zf_options* o) {
  if (!o) return;
  o->lr = 1.0;
  o->tol = 1e-5;
  o->tol_internal = 1e-12;
  o->max_iter = 1000000;
  o->max_iter_internal = 100000;
  o->max_backtrack_iter = 100;
  o->warm_start = 0;
  o->nesterov = 0;
  o->decay_rate = 0.5;
  o->nesterov_a = 0.0;
  o->nesterov_b = 0.25;
  o->deprecated = 0;
  o->dual_solver = 0;
  o->trace_capacity = 0;
  o->reserved = 0;
}
