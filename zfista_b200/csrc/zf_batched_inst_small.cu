// Explicit instantiations of the batched kernels (zf_batched_kernels.cuh): small
#include "zf_batched_kernels.cuh"

namespace zf {
template int launch_t<ZF_SD, 2, false>(const LaunchArgs&);
template int launch_t<ZF_TOI4, 2, false>(const LaunchArgs&);
template int launch_t<ZF_TOI4, 2, true>(const LaunchArgs&);
template int launch_t<ZF_TRIDIA, 3, false>(const LaunchArgs&);
template int launch_t<ZF_TRIDIA, 3, true>(const LaunchArgs&);
}  // namespace zf
