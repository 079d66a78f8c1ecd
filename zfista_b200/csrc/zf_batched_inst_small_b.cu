// Explicit instantiations of the batched kernels (zf_batched_kernels.cuh): small_b
#include "zf_batched_kernels.cuh"

namespace zf {
template int launch_t<ZF_TRIDIA, 3, 0>(const LaunchArgs&);
template int launch_t<ZF_TRIDIA, 3, 1>(const LaunchArgs&);
template int launch_t<ZF_TRIDIA, 3, 2>(const LaunchArgs&);
template int launch_t<ZF_TRIDIA, 3, 3>(const LaunchArgs&);
}  // namespace zf
