// Explicit instantiations of the batched kernels (zf_batched_kernels.cuh): lfr1b
#include "zf_batched_kernels.cuh"

namespace zf {
template int launch_t<ZF_LFR1, 3, false>(const LaunchArgs&);
template int launch_t<ZF_LFR1, 3, true>(const LaunchArgs&);
}  // namespace zf
