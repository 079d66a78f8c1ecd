// Explicit instantiations of the batched kernels (zf_batched_kernels.cuh): lfr1c
#include "zf_batched_kernels.cuh"

namespace zf {
template int launch_t<ZF_LFR1, 4, false>(const LaunchArgs&);
template int launch_t<ZF_LFR1, 4, true>(const LaunchArgs&);
}  // namespace zf
