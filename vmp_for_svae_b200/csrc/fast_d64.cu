// fast_d64.cu — instantiates the group engine of local_step_fast.cuh for D = 64 (own translation unit: the fully
// unrolled kernels take the longest to compile, one TU per D lets them build in parallel).
#define VMP_FAST_IMPL
#include "local_step_fast.cuh"

namespace vmp {
VMP_FAST_INSTANTIATE(64)
}  // namespace vmp
