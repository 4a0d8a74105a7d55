// smem_probe2.cu — wavefront cost of the D=32 engine's row reads (4 distinct 16-byte chunks shared by stride-4 lanes) for
// several row strides / layouts.  Run under ncu:
//   ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum tools/probes/smem_probe2
// one launch per pattern (1 CTA x 32 threads x ITER loads); wavefronts / ITER = cost per instruction.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 4096;

__device__ __forceinline__ int addr_of(int pat, int lane) {   // float offset
    const int lo = lane & 3, q = lane >> 2;
    switch (pat) {
        case 0: return lo * 36;            // engine today: rows of stride D+4 = 36 floats, stride-4 lanes share
        case 1: return lo * 4;             // the four chunks contiguous (64 B)
        case 2: return lo * 20;            // stride 20 floats
        case 3: return lo * 40;            // stride 40 floats (banks 0, 8, 16, 24)
        case 4: return lo * 8;             // stride 8 floats
        case 5: return (q & 3) * 36;       // same four rows, but 4 CONSECUTIVE lanes share a row
        case 6: return (q & 3) * 4;        // contiguous, consecutive lanes share
        case 7: return lo * 36 + (q >> 2) * 4 * 36;   // 8 distinct rows (halves differ)
        case 8: return q * 36;             // 8 distinct rows, consecutive lanes share
        case 9: return q * 4;              // 8 contiguous chunks, consecutive lanes share
        case 10: return lo * 4 + (q & 1) * 16;        // 8 contiguous chunks, mixed
        case 11: return (lane & 1) * 36;   // 2 distinct rows
        case 12: return (lane & 7) * 36;   // 8 distinct rows, stride-8 lanes share
        case 13: return (lane & 7) * 4;    // 8 contiguous chunks, stride-8 lanes share
        case 14: return (lane & 15) * 36;  // 16 rows stride 36
        case 15: return lane * 4;          // 32 contiguous chunks
        default: return 0;
    }
}

template <int WIDTH>
__global__ void probe(int pat, float* out, int act) {
    __shared__ __align__(16) float sm[4096];
    const int lane = threadIdx.x;
    for (int i = lane; i < 4096; i += 32) sm[i] = (float)i;
    __syncthreads();
    const unsigned a = (unsigned)__cvta_generic_to_shared(sm + addr_of(pat, lane));
    float acc = 0.f;
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
        float x, y, z, w;
        const unsigned a2 = a + ((acc == 12345.f && act) ? 16u : 0u);
        if (WIDTH == 16) { asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(a2) : "memory"); acc += x + w; }
        if (WIDTH == 8) { asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y) : "r"(a2) : "memory"); acc += x + y; }
        if (WIDTH == 4) { asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(a2) : "memory"); acc += x; }
    }
    out[lane] = acc;
}

int main() {
    float* out;
    cudaMalloc(&out, 4096);
    for (int pat = 0; pat <= 15; ++pat) probe<16><<<1, 32>>>(pat, out, 1);    // launches 0..15  LDS.128
    for (int pat = 0; pat <= 15; ++pat) probe<8><<<1, 32>>>(pat, out, 1);     // launches 16..31 LDS.64
    for (int pat = 0; pat <= 15; ++pat) probe<4><<<1, 32>>>(pat, out, 1);     // launches 32..47 LDS.32
    cudaDeviceSynchronize();
    printf("done %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
