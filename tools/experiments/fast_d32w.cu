// fast_d32w.cu — instantiates the group engine of local_step_fast.cuh for D = 32 with 16 lanes per pair (own translation
// unit: the fully unrolled kernels take the longest to compile, one TU per shape lets them build in parallel).
#define VMP_FAST_IMPL
#include "local_step_fast.cuh"

namespace vmp {
VMP_FAST_INSTANTIATE(32, 16)
}  // namespace vmp
