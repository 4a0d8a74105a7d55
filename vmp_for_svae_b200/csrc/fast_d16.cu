// fast_d16.cu — instantiates the group engine of local_step_fast.cuh for D = 16 (own translation unit: the fully
// unrolled kernels take the longest to compile, one TU per D lets them build in parallel).
#define VMP_FAST_IMPL
#include "local_step_fast.cuh"

namespace vmp {
VMP_FAST_INSTANTIATE(16)
}  // namespace vmp
