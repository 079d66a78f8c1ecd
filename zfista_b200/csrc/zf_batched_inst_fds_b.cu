// Explicit instantiations of the batched kernels (zf_batched_kernels.cuh): fds_b
#include "zf_batched_kernels.cuh"

namespace zf {
template int launch_t<ZF_FDS, 3, 1>(const LaunchArgs&);
template int launch_t<ZF_FDS, 3, 3>(const LaunchArgs&);
}  // namespace zf
