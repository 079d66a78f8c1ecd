// Explicit instantiations of the batched kernels (zf_batched_kernels.cuh): lfr1a
#include "zf_batched_kernels.cuh"

namespace zf {
template int launch_t<ZF_LFR1, 1, false>(const LaunchArgs&);
template int launch_t<ZF_LFR1, 1, true>(const LaunchArgs&);
template int launch_t<ZF_LFR1, 2, false>(const LaunchArgs&);
template int launch_t<ZF_LFR1, 2, true>(const LaunchArgs&);
}  // namespace zf
