// Explicit instantiations of the batched kernels (zf_batched_kernels.cuh): lsq
#include "zf_batched_kernels.cuh"

namespace zf {
template int launch_t<ZF_LSQ_L1, 1, false>(const LaunchArgs&);
template int launch_t<ZF_LSQ_L1, 2, false>(const LaunchArgs&);
template int launch_t<ZF_LSQ_L1, 3, false>(const LaunchArgs&);
}  // namespace zf
