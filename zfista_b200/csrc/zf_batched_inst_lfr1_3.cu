// Explicit instantiations of the batched kernels (zf_batched_kernels.cuh): lfr1_3
#include "zf_batched_kernels.cuh"

namespace zf {
template int launch_t<ZF_LFR1, 3, 0>(const LaunchArgs&);
template int launch_t<ZF_LFR1, 3, 1>(const LaunchArgs&);
template int launch_t<ZF_LFR1, 3, 2>(const LaunchArgs&);
template int launch_t<ZF_LFR1, 3, 3>(const LaunchArgs&);
}  // namespace zf
