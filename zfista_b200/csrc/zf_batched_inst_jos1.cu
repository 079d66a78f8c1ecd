// Explicit instantiations of the batched kernels (zf_batched_kernels.cuh): jos1
#include "zf_batched_kernels.cuh"

namespace zf {
template int launch_t<ZF_JOS1, 2, 0>(const LaunchArgs&);
template int launch_t<ZF_JOS1, 2, 1>(const LaunchArgs&);
template int launch_t<ZF_JOS1, 2, 2>(const LaunchArgs&);
template int launch_t<ZF_JOS1, 2, 3>(const LaunchArgs&);
}  // namespace zf
