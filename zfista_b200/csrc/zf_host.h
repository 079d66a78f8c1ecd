// Host-side helpers shared by the translation units of libzfista_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "zfista_b200.h"

namespace zf {

// record an error message for zf_last_error() and return `code`
int zf_fail(int code, const char* fmt, ...);
int zf_fail_cuda(cudaError_t e, const char* what);
// ZF_OK iff a CUDA device is usable; there is no CPU fallback behind any entry point
int zf_require_device();
// every kernel launch of this library goes through here (bench.py reports the count)
void zf_count_launch(int64_t n = 1);

}  // namespace zf

extern "C" {
// number of kernels this library has launched in this process (not in the public header's
// reference table: it has no reference counterpart, it exists for bench.py's gpu_launches)
int64_t zf_launch_count(void);
}
