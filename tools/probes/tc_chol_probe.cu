// tc_chol_probe.cu — measurements that decide the design of a tcgen05 blocked Cholesky for 64 x 64 systems (round 2):
//   P1  cycles per tcgen05.mma kind::tf32 (K = 8) for M in {64,128}, N in {16,32,48,64}, issued back to back into ONE
//       accumulator (dependent chain) or round-robin into 8 accumulators (independent)
//   P2  an M = 64 accumulator placed at TMEM lane offset 16 (two 64-row systems interleaved in one 128-lane block) and
//       the negate-A bit (D -= A B^T) with A and B descriptors on the SAME shared-memory tile (L21 L21^T)
//   P3  tcgen05.ld 32x32b.x16 cycles per load (one warp / four warps)
//   P4  round trip: 6 MMAs (M=64, N=48) -> commit -> mbarrier wait -> tcgen05.ld of 16 columns, in cycles
//   P5  legacy mma.sync.m16n8k8 tf32 rate (whole GPU)
// Every wait is bounded; a wrong descriptor cannot hang the box.
// Build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/probes/tc_chol_probe tools/probes/tc_chol_probe.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint32_t idesc_tf32(int M, int N, bool negA) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (negA ? (1u << 13) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t taddr, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(taddr), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool wait_bar(uint64_t* bar, uint32_t parity) {
    for (int spin = 0; spin < (1 << 20); ++spin) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) return true;
    }
    return false;
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}

constexpr int KT = 16;                       // K of the staged tiles (two tf32 MMA steps)
constexpr uint32_t LBO = 128, SBO = (KT / 4) * 128;
__device__ __forceinline__ int tile_off(int m, int k) { return ((m / 8) * SBO + (k / 4) * LBO + (m % 8) * 16 + (k % 4) * 4) / 4; }

struct Out {
    long long mma_cycles[2][4][2];   // [M64|M128][N16..64][dependent|independent]   cycles for REPS MMAs
    long long ld_cycles[2];          // one warp | four warps, cycles for REPS loads of x16
    long long roundtrip;             // cycles for REPS round trips
    int status;
};
constexpr int REPS = 256;

__global__ void __launch_bounds__(128) timing_probe(Out* out) {
    __shared__ __align__(1024) float sA[128 * KT];
    __shared__ __align__(1024) float sB[64 * KT];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < 128 * KT; e += 128) sA[e] = 1.0f + 0.001f * (e % 7);
    for (int e = tid; e < 64 * KT; e += 128) sB[e] = 0.5f + 0.002f * (e % 5);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem_base;
    uint32_t parity = 0;
    bool ok = true;
    const uint64_t da = make_desc(smem_u32(sA), LBO, SBO), db = make_desc(smem_u32(sB), LBO, SBO);

    // ---- P1
    for (int mi = 0; mi < 2; ++mi)
        for (int ni = 0; ni < 4; ++ni)
            for (int ind = 0; ind < 2; ++ind) {
                const int M = mi ? 128 : 64, N = 16 * (ni + 1);
                const uint32_t idesc = idesc_tf32(M, N, false);
                __syncthreads();
                long long t0 = 0, t1 = 0;
                if (tid == 0) {
                    t0 = clock64();
                    for (int r = 0; r < REPS; ++r) {
                        const uint32_t col = ind ? (uint32_t)((r & 7) * 64) : 0u;
                        mma_tf32(taddr + col, da, db, idesc, r >= 8 ? 1u : 0u);
                    }
                    commit(&bar);
                }
                ok = wait_bar(&bar, parity) && ok;
                parity ^= 1;
                if (tid == 0) {
                    t1 = clock64();
                    out->mma_cycles[mi][ni][ind] = t1 - t0;
                }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
    // ---- P3
    for (int four = 0; four < 2; ++four) {
        __syncthreads();
        uint32_t v[16];
        uint32_t sink = 0;
        const long long t0 = clock64();
        if (four || warp == 0) {
            for (int r = 0; r < REPS; ++r) {
                ld16(taddr + ((uint32_t)(warp * 32) << 16) + (uint32_t)((r & 31) * 16), v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int i = 0; i < 16; ++i) sink ^= v[i];
            }
        }
        const long long t1 = clock64();
        if (sink == 0x12345678u) out->status = 77;
        __syncthreads();
        if (tid == 0) out->ld_cycles[four] = t1 - t0;
    }
    // ---- P4: round trip
    {
        const uint32_t idesc = idesc_tf32(64, 48, true);
        __syncthreads();
        uint32_t sink = 0;
        const long long t0 = clock64();
        for (int r = 0; r < REPS; ++r) {
            if (tid == 0) {
#pragma unroll
                for (int q = 0; q < 6; ++q) mma_tf32(taddr, da + (uint64_t)((q & 1) * 256 >> 4), da + (uint64_t)((q & 1) * 256 >> 4), idesc, 1u);
                commit(&bar);
            }
            ok = wait_bar(&bar, parity) && ok;
            parity ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t v[16];
            ld16(taddr + ((uint32_t)(warp * 32) << 16), v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            sink ^= v[0];
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
        }
        const long long t1 = clock64();
        if (sink == 0x12345678u) out->status = 78;
        if (tid == 0) out->roundtrip = t1 - t0;
    }
    if (tid == 0) out->status = ok ? 1 : -1;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(taddr) : "memory");
}

// ---- P2: two 64-row systems interleaved (lane offsets 0 and 16), D = C - L L^T with A and B on the same tile
__global__ void __launch_bounds__(128) layout_probe(const float* __restrict__ L0, const float* __restrict__ L1,
                                                    float* __restrict__ out, int* __restrict__ status) {
    __shared__ __align__(1024) float sL[2][64 * KT];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < 64 * KT; e += 128) {
        sL[0][tile_off(e / KT, e % KT)] = L0[e];
        sL[1][tile_off(e / KT, e % KT)] = L1[e];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem_base;
    // initialise the accumulators with a known C through tcgen05.st: C[row][col] = row + 0.01 col + 1000 pair
    {
        const int pair = lane >> 4, row = warp * 16 + (lane & 15);
        uint32_t v[16];
        for (int c0 = 0; c0 < 64; c0 += 16) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __float_as_uint((float)row + 0.01f * (float)(c0 + i) + 1000.f * (float)pair);
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                         ::"r"(taddr + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                           "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]),
                           "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        const uint32_t idesc = idesc_tf32(64, 64, true);           // negate A: D = D - L L^T
        for (int pair = 0; pair < 2; ++pair)
            for (int ks = 0; ks < KT / 8; ++ks) {
                const uint64_t d = make_desc(smem_u32(sL[pair]) + ks * 2 * LBO, LBO, SBO);
                mma_tf32(taddr + ((uint32_t)(16 * pair) << 16), d, d, idesc, 1u);
            }
        commit(&bar);
    }
    const bool ok = wait_bar(&bar, 0);
    if (tid == 0) status[0] = ok ? 1 : -1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (ok) {
        uint32_t v[16];
        for (int c0 = 0; c0 < 64; c0 += 16) {
            ld16(taddr + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * 64 + c0 + i] = __uint_as_float(v[i]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(taddr) : "memory");
}

// ---- P5: legacy mma.sync tf32
__global__ void __launch_bounds__(256) mma_sync_probe(float* out, int iters) {
    float c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    uint32_t a[4] = {0x3f800000u + threadIdx.x, 0x3f801000u, 0x3f802000u, 0x3f803000u}, b[2] = {0x3f000000u, 0x3f004000u};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static float tf32_trunc(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u &= 0xFFFFE000u;
    memcpy(&x, &u, 4);
    return x;
}

int main() {
    // P1 / P3 / P4
    Out* dout;
    cudaMalloc(&dout, sizeof(Out));
    cudaMemset(dout, 0, sizeof(Out));
    timing_probe<<<1, 128>>>(dout);
    cudaError_t e = cudaDeviceSynchronize();
    Out h;
    cudaMemcpy(&h, dout, sizeof(Out), cudaMemcpyDeviceToHost);
    printf("timing_probe: cuda=%s status=%d\n", cudaGetErrorString(e), h.status);
    if (e == cudaSuccess && h.status == 1) {
        for (int mi = 0; mi < 2; ++mi)
            for (int ni = 0; ni < 4; ++ni)
                printf("P1 M=%3d N=%2d K=8 tf32: %7.1f cycles/MMA dependent (one accumulator), %7.1f independent (8 accumulators); "
                       "floor max(M,128)*N/256 = %d\n", mi ? 128 : 64, 16 * (ni + 1), (double)h.mma_cycles[mi][ni][0] / REPS,
                       (double)h.mma_cycles[mi][ni][1] / REPS, 128 * 16 * (ni + 1) / 256);
        printf("P3 tcgen05.ld 32x32b.x16 + wait: %.1f cycles/load with one warp, %.1f with four warps (each warp its own 32 lanes)\n",
               (double)h.ld_cycles[0] / REPS, (double)h.ld_cycles[1] / REPS);
        printf("P4 round trip (6 MMAs M=64 N=48 K=8 -> commit -> mbarrier -> ld x16 -> __syncthreads): %.1f cycles\n",
               (double)h.roundtrip / REPS);
    }
    // P2
    {
        std::vector<float> L0(64 * KT), L1(64 * KT), out(128 * 64, -777.f);
        srand(3);
        for (auto& v : L0) v = (float)(rand() % 2001 - 1000) / 512.f;
        for (auto& v : L1) v = (float)(rand() % 2001 - 1000) / 512.f;
        float *d0, *d1, *dO;
        int* dS;
        cudaMalloc(&d0, L0.size() * 4); cudaMalloc(&d1, L1.size() * 4); cudaMalloc(&dO, out.size() * 4); cudaMalloc(&dS, 4);
        cudaMemcpy(d0, L0.data(), L0.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(d1, L1.data(), L1.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dO, out.data(), out.size() * 4, cudaMemcpyHostToDevice);
        cudaMemset(dS, 0, 4);
        layout_probe<<<1, 128>>>(d0, d1, dO, dS);
        e = cudaDeviceSynchronize();
        int st = 0;
        cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost);
        printf("layout_probe: cuda=%s status=%d\n", cudaGetErrorString(e), st);
        if (e == cudaSuccess && st == 1) {
            double worst = 0;
            int bad = 0;
            for (int lanei = 0; lanei < 128; ++lanei) {
                const int pair = (lanei & 31) >> 4, row = (lanei >> 5) * 16 + (lanei & 15);
                const std::vector<float>& L = pair ? L1 : L0;
                for (int c = 0; c < 64; ++c) {
                    double s = (double)((float)row + 0.01f * (float)c + 1000.f * (float)pair);
                    for (int k = 0; k < KT; ++k) s -= (double)tf32_trunc(L[row * KT + k]) * (double)tf32_trunc(L[c * KT + k]);
                    const double err = fabs(out[lanei * 64 + c] - s);
                    worst = fmax(worst, err);
                    if (err > 2e-3) ++bad;
                }
            }
            printf("P2 two interleaved M=64 accumulators (lane offsets 0/16), D = C - L L^T via negate-A, A and B on one tile: "
                   "%d mismatches, worst abs err %.3e\n", bad, worst);
            if (bad) {
                printf("   sample: lane 0 col 0..3: %g %g %g %g ; lane 16 col 0..3: %g %g %g %g\n", out[0], out[1], out[2], out[3],
                       out[16 * 64], out[16 * 64 + 1], out[16 * 64 + 2], out[16 * 64 + 3]);
            }
        }
    }
    // P5
    {
        float* dO;
        const int grid = 148 * 8, iters = 4096;
        cudaMalloc(&dO, (size_t)grid * 256 * 4);
        mma_sync_probe<<<grid, 256>>>(dO, 16);
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        mma_sync_probe<<<grid, 256>>>(dO, iters);
        cudaEventRecord(b);
        e = cudaDeviceSynchronize();
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        const double flops = (double)grid * 8 /*warps*/ * iters * 8 * (2.0 * 16 * 8 * 8);
        printf("P5 mma.sync.m16n8k8 tf32: cuda=%s %.3f ms -> %.1f TFLOP/s dense (%.2f cycles per MMA per SM sub-partition at 1.9 GHz)\n",
               cudaGetErrorString(e), ms, flops / (ms * 1e-3) / 1e12,
               (ms * 1e-3 * 1.9e9) / ((double)grid * 8 * iters * 8 / (148.0 * 4)));
    }
    return 0;
}
