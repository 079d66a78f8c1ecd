// Explicit instantiations of the batched kernels (zf_batched_kernels.cuh): fds_a
#include "zf_batched_kernels.cuh"

namespace zf {
template int launch_t<ZF_FDS, 3, 0>(const LaunchArgs&);
template int launch_t<ZF_FDS, 3, 2>(const LaunchArgs&);
}  // namespace zf
