// tcgen05_probe.cu — minimal, self-checking tcgen05 (5th-gen tensor core) TF32 MMA on sm_100a, written to pin down the
// shared-memory / instruction descriptor encodings and the TMEM accumulator layouts on THIS hardware before the
// round-2 engine work (blocked Cholesky trailing updates and the X^T diag(r) X statistics on the tensor pipe).
//
//   D[M x N] (fp32, TMEM) = A[M x 16] * B[N x 16]^T,  A / B tf32 in shared memory, K-major, no swizzle,
//   two tcgen05.mma (K = 8 each, the second accumulating), one tcgen05.commit -> mbarrier, tcgen05.ld back.
//
// Layout used (canonical K-major "interleave"/no-swizzle form): 8-row x 16-byte core matrices, 128 B each;
//   addr(m, k) = (m / 8) * SBO + (k / 4) * LBO + (m % 8) * 16 + (k % 4) * 4,  LBO = 128 B, SBO = (K/4) * 128 B.
// Shared-memory descriptor (64 bit): [0,14) addr>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version = 1 | [61,64) swizzle = 0.
// Instruction descriptor (32 bit): [4,6) D fmt (1 = f32) | [7,10) A fmt (2 = tf32) | [10,13) B fmt | [15] A major (0 = K) |
//   [16] B major | [17,23) N >> 3 | [24,29) M >> 4.
// Build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/probes/tcgen05_probe tools/probes/tcgen05_probe.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int KTOT = 16, NCOL = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                                  // descriptor version (Blackwell)
    return d;
}

template <int M>
__global__ void __launch_bounds__(128) probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ out,
                                             int* __restrict__ status) {
    __shared__ __align__(128) float sA[128 * KTOT];
    __shared__ __align__(128) float sB[NCOL * KTOT];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr uint32_t LBO = 128, SBO = (KTOT / 4) * 128;

    for (int e = tid; e < M * KTOT; e += 128) {
        const int m = e / KTOT, k = e % KTOT;
        sA[((m / 8) * SBO + (k / 4) * LBO + (m % 8) * 16 + (k % 4) * 4) / 4] = A[e];
    }
    for (int e = tid; e < NCOL * KTOT; e += 128) {
        const int n = e / KTOT, k = e % KTOT;
        sB[((n / 8) * SBO + (k / 4) * LBO + (n % 8) * 16 + (k % 4) * 4) / 4] = B[e];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem_base;

    if (tid == 0) {
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NCOL >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
#pragma unroll
        for (int ks = 0; ks < KTOT / 8; ++ks) {
            const uint64_t da = make_desc(smem_u32(sA) + ks * 2 * LBO, LBO, SBO);
            const uint64_t db = make_desc(smem_u32(sB) + ks * 2 * LBO, LBO, SBO);
            const uint32_t acc = ks > 0 ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(taddr), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // bounded wait: a wrong descriptor must not hang the box
    bool ok = false;
    for (int spin = 0; spin < (1 << 22) && !ok; ++spin) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
        ok = done != 0;
    }
    if (tid == 0) status[0] = ok ? 1 : -1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (ok) {
        uint32_t v[NCOL];
        const uint32_t a = taddr + ((uint32_t)(warp * 32) << 16);
#pragma unroll
        for (int c = 0; c < NCOL; c += 8) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(v[c]), "=r"(v[c + 1]), "=r"(v[c + 2]), "=r"(v[c + 3]), "=r"(v[c + 4]), "=r"(v[c + 5]),
                           "=r"(v[c + 6]), "=r"(v[c + 7])
                         : "r"(a + c));
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < NCOL; ++c) out[(warp * 32 + lane) * NCOL + c] = __uint_as_float(v[c]);   // row = TMEM lane
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(taddr) : "memory");
}

static float tf32(float x) {           // round-to-nearest-even to 10 mantissa bits (what the tensor core consumes: it truncates)
    uint32_t u;
    memcpy(&u, &x, 4);
    u &= 0xFFFFE000u;                  // hardware reads the upper 19 bits of the fp32 container
    memcpy(&x, &u, 4);
    return x;
}

template <int M>
static int run() {
    std::vector<float> A(128 * KTOT, 0.f), B(NCOL * KTOT), out(128 * NCOL, -777.f);
    srand(7 + M);
    for (int i = 0; i < M * KTOT; ++i) A[i] = (float)(rand() % 2001 - 1000) / 512.f;
    for (auto& v : B) v = (float)(rand() % 2001 - 1000) / 512.f;
    float *dA, *dB, *dO;
    int* dS;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dO, out.size() * 4); cudaMalloc(&dS, 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dO, out.data(), out.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dS, 0, 4);
    probe<M><<<1, 128>>>(dA, dB, dO, dS);
    cudaError_t e = cudaDeviceSynchronize();
    int st = 0;
    cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost);
    printf("M=%d: cuda=%s status=%d\n", M, cudaGetErrorString(e), st);
    if (e != cudaSuccess || st != 1) return 1;
    // which TMEM lane holds which row?  match every lane's 64 outputs against every reference row
    std::vector<double> ref(M * NCOL);
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < NCOL; ++n) {
            double s = 0;
            for (int k = 0; k < KTOT; ++k) s += (double)tf32(A[m * KTOT + k]) * (double)tf32(B[n * KTOT + k]);
            ref[m * NCOL + n] = s;
        }
    int matched = 0;
    double worst = 0;
    for (int l = 0; l < 128; ++l) {
        int best = -1;
        double berr = 1e30;
        for (int m = 0; m < M; ++m) {
            double err = 0;
            for (int n = 0; n < NCOL; ++n) err = fmax(err, fabs(out[l * NCOL + n] - ref[m * NCOL + n]));
            if (err < berr) { berr = err; best = m; }
        }
        if (berr < 1e-3) {
            ++matched;
            worst = fmax(worst, berr);
            if (l < 4 || l % 16 == 0 || best != l) printf("  lane %3d <- row %3d  (max err %.2e)\n", l, best, berr);
        } else if (l < 4 || l % 16 == 0) {
            printf("  lane %3d: no row matches (first value %g)\n", l, out[l * NCOL]);
        }
    }
    printf("M=%d: %d lanes hold a row of A B^T, worst abs err %.3e\n", M, matched, worst);
    return matched == M ? 0 : 2;
}

int main() {
    int rc = run<128>();
    rc |= run<64>() << 4;
    printf("probe rc=%d\n", rc);
    return 0;
}
