// Explicit instantiations of the batched kernels (zf_batched_kernels.cuh): small_a
#include "zf_batched_kernels.cuh"

namespace zf {
template int launch_t<ZF_SD, 2, 2>(const LaunchArgs&);
template int launch_t<ZF_ZDT1, 2, 2>(const LaunchArgs&);
template int launch_t<ZF_TOI4, 2, 0>(const LaunchArgs&);
template int launch_t<ZF_TOI4, 2, 1>(const LaunchArgs&);
template int launch_t<ZF_TOI4, 2, 2>(const LaunchArgs&);
template int launch_t<ZF_TOI4, 2, 3>(const LaunchArgs&);
}  // namespace zf
