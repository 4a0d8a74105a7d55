// fast_d16w.cu — instantiates the group engine of local_step_fast.cuh for D = 16 with 8 lanes per pair (own translation
// unit: the fully unrolled kernels take the longest to compile, one TU per shape lets them build in parallel).
#define VMP_FAST_IMPL
#include "local_step_fast.cuh"

namespace vmp {
VMP_FAST_INSTANTIATE(16, 8)
}  // namespace vmp
