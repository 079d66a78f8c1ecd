// Explicit instantiations of the batched kernels (zf_batched_kernels.cuh): lfr1_4
#include "zf_batched_kernels.cuh"

namespace zf {
template int launch_t<ZF_LFR1, 4, 0>(const LaunchArgs&);
template int launch_t<ZF_LFR1, 4, 1>(const LaunchArgs&);
template int launch_t<ZF_LFR1, 4, 2>(const LaunchArgs&);
template int launch_t<ZF_LFR1, 4, 3>(const LaunchArgs&);
}  // namespace zf
