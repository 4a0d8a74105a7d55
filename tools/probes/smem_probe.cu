// smem_probe.cu — shared-memory wavefront cost of the access patterns the group engines use (run under ncu:
//   ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum tools/probes/smem_probe
// each pattern is one launch of 1 CTA x 32 threads x ITER instructions; wavefronts / ITER = cost per instruction).
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 4096;
constexpr int RS = 20, GS = 720;

__device__ __forceinline__ int addr_of(int pat, int lane) {   // float offset
    const int g = lane >> 4, gl = lane & 15;
    const int hi = gl >> 2, lo = gl & 3;
    switch (pat) {
        case 0: return 0;                                   // all lanes same address
        case 1: return g * GS;                              // half-warps distinct (old engine column broadcast)
        case 2: return g * GS + hi * RS;                    // address by the high 2 bits of gl (4 lanes in a row share)
        case 3: return g * GS + lo * RS;                    // address by the low 2 bits of gl (stride-4 lanes share)
        case 4: return g * GS + ((hi + lo) & 3) * RS;       // skewed classes
        case 5: return gl * 4;                              // 16 contiguous chunks, halves identical (lane-major record)
        case 6: return gl * 68;                             // 16 rows, stride 68 floats, halves identical (old P2 rows)
        case 7: return g * GS + gl * 4;                     // 32 distinct contiguous-per-half chunks
        case 8: return (g * 2 + (gl >> 3)) * 160;           // quarter-warps distinct, same within quarter
        case 9: return g * GS + (hi & 1) * RS + (hi >> 1) * 2 * RS;   // same as 2
        default: return 0;
    }
}

template <int WIDTH, bool STORE>
__global__ void probe(int pat, int active_mask_kind, float* out, int base = 0) {
    __shared__ __align__(16) float sm[4096];
    const int lane = threadIdx.x;
    for (int i = lane; i < 4096; i += 32) sm[i] = (float)i;
    __syncthreads();
    const int gl = lane & 15;
    bool act = true;
    if (active_mask_kind == 1) act = (gl & 3) == 1;          // stride-4 lanes (publish with gl = pr*4+pc)
    if (active_mask_kind == 2) act = (gl >> 2) == 1;         // 4 consecutive lanes (publish with gl = pc*4+pr)
    const unsigned a = (unsigned)__cvta_generic_to_shared(sm + base + addr_of(pat, lane));
    float acc = 0.f;
    for (int it = 0; it < ITER; ++it) {
        if (act) {
            if (STORE) {
                if (WIDTH == 16) asm volatile("st.shared.v4.f32 [%0], {%1,%1,%1,%1};" ::"r"(a), "f"(acc));
                if (WIDTH == 8) asm volatile("st.shared.v2.f32 [%0], {%1,%1};" ::"r"(a), "f"(acc));
                if (WIDTH == 4) asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(acc));
            } else {
                float x, y, z, w;
                if (WIDTH == 16) { asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(a)); acc += x + w; }
                if (WIDTH == 8) { asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y) : "r"(a)); acc += x + y; }
                if (WIDTH == 4) { asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(a)); acc += x; }
            }
        }
    }
    out[lane] = acc + sm[lane];
}

int main() {
    float* out;
    cudaMalloc(&out, 4096);
    for (int pat = 0; pat <= 8; ++pat) { probe<16, false><<<1, 32>>>(pat, 0, out); }      // launches 0..8   LDS.128
    for (int pat = 0; pat <= 8; ++pat) { probe<8, false><<<1, 32>>>(pat, 0, out); }       // launches 9..17  LDS.64
    for (int pat = 0; pat <= 8; ++pat) { probe<4, false><<<1, 32>>>(pat, 0, out); }       // launches 18..26 LDS.32
    for (int k = 1; k <= 2; ++k) {
        probe<16, true><<<1, 32>>>(2, k, out);                                              // 27,30  STS.128 by hi-class, masks
        probe<8, true><<<1, 32>>>(2, k, out);                                               // 28,31
        probe<4, true><<<1, 32>>>(2, k, out);                                               // 29,32
    }
    probe<16, true><<<1, 32>>>(7, 0, out);                                                  // 33 STS.128 all lanes distinct
    for (int base = 4; base <= 28; base += 4) { probe<16, false><<<1, 32>>>(2, 0, out, base); probe<16, false><<<1, 32>>>(3, 0, out, base); }  // 34..47
    for (int k = 0; k <= 2; ++k) { probe<16, true><<<1, 32>>>(2, k, out, 16); probe<8, true><<<1, 32>>>(2, k, out, 16); probe<4, true><<<1, 32>>>(2, k, out, 16); }   // 48..56
    cudaDeviceSynchronize();
    printf("done %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
