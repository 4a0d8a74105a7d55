// smem_probe3.cu — shared-memory wavefront cost of LDS.128 / LDS.64 issued in BURSTS of independent loads (as the engines
// issue them) for the lane->address patterns of the D=32 engine and candidate re-mappings.  Run under
//   ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum tools/probes/smem_probe3
// (a loop of one dependent load reports 2 wavefronts for patterns that cost 4 in a burst: tools/probes/smem_probe2.cu)
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITER = 64;
__device__ __forceinline__ int addr_of(int pat, int lane) {   // float offset: class = the lane bits selected by mask `pat`, compacted
    int cls = 0, nb = 0;
    for (int b = 0; b < 5; ++b)
        if (pat >> b & 1) { cls |= ((lane >> b) & 1) << nb; ++nb; }
    return cls * 36;                  // rows of 36 floats: class c sits in banks 4c..4c+3 (mod 32)
}
template <int WIDTH>
__global__ void probe(int pat, float* out, int zero) {
    extern __shared__ __align__(128) float sm[];
    const int tid = threadIdx.x;
    for (int i = tid; i < 8192; i += blockDim.x) sm[i] = (float)i;
    __syncthreads();
    const float* base = sm + addr_of(pat, tid & 31);
    float acc = 0.f;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int j = 0; j < 64; ++j) {
            if (WIDTH == 16) {
                const float4 v = reinterpret_cast<const float4*>(base)[j];
                acc += v.x * v.y + v.z * v.w;
            } else {
                const float2 v = reinterpret_cast<const float2*>(base)[j];
                acc += v.x * v.y;
            }
        }
        base += zero;
    }
    out[tid] = acc;
}
int main() {
    float* out;
    cudaMalloc(&out, 4096);
    cudaFuncSetAttribute(probe<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(probe<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int pat = 0; pat < 32; ++pat) probe<16><<<1, 32, 65536>>>(pat, out, 0);
    for (int pat = 0; pat < 32; ++pat) probe<8><<<1, 32, 65536>>>(pat, out, 0);
    cudaDeviceSynchronize();
    printf("done %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
