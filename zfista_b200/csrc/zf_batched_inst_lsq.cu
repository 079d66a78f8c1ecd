// Explicit instantiations of the batched kernels (zf_batched_kernels.cuh): lsq
#include "zf_batched_kernels.cuh"

namespace zf {
template int launch_t<ZF_LSQ_L1, 1, 0>(const LaunchArgs&);
template int launch_t<ZF_LSQ_L1, 2, 0>(const LaunchArgs&);
template int launch_t<ZF_LSQ_L1, 3, 0>(const LaunchArgs&);
}  // namespace zf
