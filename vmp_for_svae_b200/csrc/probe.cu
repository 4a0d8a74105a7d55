// probe.cu — FP32 pipe micro-benchmark: sustained FMA throughput of scalar FFMA vs packed FFMA2 (fma.rn.f32x2).
// bench.py reports the local-step roofline against the nominal FP32 peak; this probe measures what the part
// actually sustains for both instruction forms (the roofline denominator needs FFMA2 to be reachable at all).
#include "common.cuh"

namespace vmp {

template <bool PACKED>
__global__ void __launch_bounds__(256) fma_probe_kernel(int iters, float seed, float* out) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + (float)(threadIdx.x + i);
    const float m0 = 1.0f + seed * 1e-7f, m1 = 1.0f - seed * 1e-7f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep) {
            if (PACKED) {
#pragma unroll
                for (int i = 0; i < 16; i += 2)
                    asm volatile(
                        "{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2,%2};\n\tmov.b64 rb, {%3,%3};\n\tmov.b64 rc, {%0,%1};\n\t"
                        "fma.rn.f32x2 rc, rc, ra, rb;\n\tmov.b64 {%0,%1}, rc;\n\t}"
                        : "+f"(a[i]), "+f"(a[i + 1])
                        : "f"(m0), "f"(m1));
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], m0, m1);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace vmp

extern "C" {
// Launches grid x 256 threads, each doing iters * 128 FMAs; returns 0 and the caller times it with CUDA events.
int vmp_fma_probe(int packed, int grid, int iters, float* out, void* stream) {
    if (grid <= 0 || iters <= 0 || !out) return VMP_E_BADARG;
    if (packed) vmp::fma_probe_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(iters, 1.0f, out);
    else vmp::fma_probe_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(iters, 1.0f, out);
    return vmp::launch_status();
}
}
