// Explicit instantiations of the batched kernels (zf_batched_kernels.cuh): jos1
#include "zf_batched_kernels.cuh"

namespace zf {
template int launch_t<ZF_JOS1, 2, false>(const LaunchArgs&);
template int launch_t<ZF_JOS1, 2, true>(const LaunchArgs&);
template int launch_t<ZF_ZDT1, 2, false>(const LaunchArgs&);
}  // namespace zf
