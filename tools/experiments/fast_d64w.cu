// fast_d64w.cu — instantiates the group engine of local_step_fast.cuh for D = 64 with 32 lanes per pair (own translation
// unit: the fully unrolled kernels take the longest to compile, one TU per shape lets them build in parallel).
#define VMP_FAST_IMPL
#include "local_step_fast.cuh"

namespace vmp {
VMP_FAST_INSTANTIATE(64, 32)
}  // namespace vmp
