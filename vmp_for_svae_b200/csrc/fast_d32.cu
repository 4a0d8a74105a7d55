// fast_d32.cu — instantiates the group engine of local_step_fast.cuh for D = 32 (own translation unit: the fully unrolled
// kernels take the longest to compile, one TU per shape lets them build in parallel).  VMP_D32_LANES lanes per pair:
// 8 (4 rows per lane, 128 registers, two CTAs per SM) or 4 (8 rows per lane, 255 registers: every broadcast column read
// serves 8 pairs instead of 4 — half the shared-memory wavefronts per pair).
#define VMP_FAST_IMPL
#include "local_step_fast.cuh"

namespace vmp {
VMP_FAST_INSTANTIATE(32, VMP_D32_LANES)
}  // namespace vmp
