// fast2d_d64.cu — instantiates the 2-D cyclic group engine (local_step_fast2d.cuh) for D = 64 and its record packer.
#define VMP_FAST_IMPL
#define VMP_FAST2D_IMPL
#include "local_step_fast2d.cuh"
