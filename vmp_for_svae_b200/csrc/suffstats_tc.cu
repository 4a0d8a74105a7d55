// suffstats_tc.cu — the responsibility-weighted second-moment contraction  S_k = sum_n w_nk x_n x_n^T  on the 5th-gen
// tensor cores (tcgen05, accumulators in TMEM) for fp32, D = 64 (north_star: "tensor cores only for the X^T diag(r_k) X
// contraction and only if it stays within tolerance").  Replaces the reductions of gmm.m_step / svae.m_step
// (gmm.py:25-46, 201-227; svae.py:154-176) for the C5 shape; every other shape keeps the FP32 kernels of suffstats.cu.
//
// fp32 accuracy on a tf32 pipe: split-operand products.  The tensor core reads the upper 19 bits of each fp32 container
// (verified: tools/probes/tcgen05_probe.cu); with hi(v) = rna_tf32(v) and lo(v) = v - hi(v) (exact)
//   a b  ~=  hi(a) hi(b) + lo(a) hi(b) + hi(a) lo(b)            (dropped term lo lo ~ 2^-22 relative)
// three MMAs per K-step.  The TMEM accumulator is fp32; it is drained into fp64 shared-memory accumulators every
// TC_FLUSH chunks (512 points) so that long sums keep double accuracy, like the FP32 kernel's register runs (the
// tensor core's fp32 accumulation truncates; measured worst case 3e-6 of a block's magnitude at 1024-point runs).
//
// One CTA = two components x a slice of points.  Per chunk of 32 points the CTA
//   1. stages the x rows and the two weight columns (cp.async, one chunk ahead),
//   2. builds six K-major operand tiles in shared memory (B = x^T hi/lo shared by both components; A_k = (w_k . x)^T hi/lo),
//      each thread also keeping its running sum_n w x_j / sum_n w,
//   3. a dedicated issuer warp launches 4 x 3 tcgen05.mma (M = 128: both components stacked, N = 64, K = 8, kind::tf32)
//      and commits them to an mbarrier;
// tile sets are double-buffered, so the tensor pipe works on chunk c while the CUDA cores build chunk c+1.
// Operand tile layout (canonical K-major, no swizzle): 8-row x 16-byte core matrices,
//   offset(row j, point n) = (j/8) * 1024 + (n/4) * 128 + (j%8) * 16 + (n%4) * 4 bytes   (LBO = 128 B, SBO = 1024 B).
// TMEM accumulator layout (measured): M = 128 puts row m in lane m; M = 64 would use lanes 32 * (i / 16) + i % 16.
#include "common.cuh"
#include "ng_tail.cuh"

namespace vmp {


constexpr int TC_D = 64, TC_PC = 32, TC_THREADS = 256, TC_FLUSH = 16;
constexpr int TC_ITEMS = (TC_D * TC_PC / 4) / TC_THREADS;      // (feature, point-quad) items per builder thread
constexpr int TC_TILE = TC_D * TC_PC;                 // floats per operand tile (8 KB)
constexpr int TC_RAWLD = TC_D + 4;                    // padded row of the staged x chunk
constexpr int TC_DLD = TC_D + 1;                      // padded row of the fp64 accumulators

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr) {          // K-major, no swizzle, LBO 128 B, SBO 1024 B
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46);
}
__device__ __forceinline__ void tc_mma(uint32_t taddr, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(taddr), "l"(da), "l"(db), "r"(idesc),
        "r"(accumulate)
        : "memory");
}
// returns false if the phase did not complete within the spin budget (a broken descriptor must not hang the GPU)
__device__ __forceinline__ bool tc_wait(uint64_t* bar, uint32_t parity) {
    for (int spin = 0; spin < (1 << 22); ++spin) {       // each try_wait blocks for a bounded time itself
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(tc_smem_u32(bar)), "r"(parity)
                     : "memory");
        if (done) return true;
    }
    return false;
}

// round-to-nearest tf32 split: hi = rna_tf32(v), lo = v - hi (exact); the hardware truncates lo's low 13 bits, an error of
// 2^-21 |v| whose sign follows lo's (random), so it does not accumulate as a bias
__device__ __forceinline__ float tc_hi(float v) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    return __uint_as_float(u);
}
__device__ __forceinline__ void tc_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void tc_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// Threads 0..TC_THREADS-1 build operand tiles and drain accumulators; the last warp only issues the MMAs (warp specialisation keeps the
// single-thread issue sequence off the builders' critical path).  The two components of the CTA are stacked along M:
// A = [(w_k0 . x)^T ; (w_k1 . x)^T] is 128 x 32 per chunk, so one M = 128 MMA serves both at the full tensor rate and
// TMEM lane m holds row m % 64 of component k0 + m / 64.
__global__ void __launch_bounds__(TC_THREADS + 32, 1)
suffstats_tc_kernel(int64_t N, int K, int64_t pts_per_slice, const float* __restrict__ x, const float* __restrict__ r,
                    int r_is_log, double* __restrict__ stats, const NgTail tail) {
    extern __shared__ __align__(1024) unsigned char smraw[];
    float* tiles = reinterpret_cast<float*>(smraw);                       // [2][6][TC_TILE]: Ahi (2 tiles) Alo (2) Bhi Blo
    float* raw = tiles + 2 * 6 * TC_TILE;                                 // [2][TC_PC][TC_RAWLD]
    float* wraw = raw + 2 * TC_PC * TC_RAWLD;                             // [2][TC_PC][2]  staged r / log r
    float* wsm = wraw + 2 * TC_PC * 2;                                    // [TC_PC][2]     weights of the current chunk
    double* dacc = reinterpret_cast<double*>(wsm + TC_PC * 2);            // [2][TC_D][TC_DLD]
    double* dsum = dacc + 2 * TC_D * TC_DLD;                              // [2][TC_D + 2]: sum w x_j | sum w
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_base;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k0 = 2 * blockIdx.x;
    const int64_t n_begin = (int64_t)blockIdx.y * pts_per_slice;
    const int64_t n_end = min(N, n_begin + pts_per_slice);
    if (n_begin >= n_end) {
        ng_tail_run<float>(tail, K, TC_D, stats, gridDim.x * gridDim.y);
        return;
    }
    const int nchunks = (int)((n_end - n_begin + TC_PC - 1) / TC_PC);

    for (int e = tid; e < 2 * TC_D * TC_DLD + 2 * (TC_D + 2); e += TC_THREADS + 32) dacc[e] = 0.0;   // dacc | dsum contiguous
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc_smem_u32(&bars[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc_smem_u32(&bars[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(tc_smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem_base;
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_D >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

    if (warp == TC_THREADS / 32) {
        // ------------------------------------------------ MMA issuer warp
        uint64_t dAhi[2], dAlo[2], dBhi[2], dBlo[2];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const uint32_t tb = tc_smem_u32(tiles + b * 6 * TC_TILE);
            dAhi[b] = tc_desc(tb);
            dAlo[b] = tc_desc(tb + 2 * TC_TILE * 4);
            dBhi[b] = tc_desc(tb + 4 * TC_TILE * 4);
            dBlo[b] = tc_desc(tb + 5 * TC_TILE * 4);
        }
        for (int it = 0; it < nchunks; ++it) {
            const int buf = it & 1;
            tc_bar_sync(2 + buf, TC_THREADS + 32);                       // tile set `buf` is complete and fenced
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const bool fresh = (it % TC_FLUSH) == 0;                  // first chunk after a drain overwrites TMEM
#pragma unroll
                for (int ks = 0; ks < TC_PC / 8; ++ks) {
                    const uint64_t o = (uint64_t)(ks * 256 >> 4);         // two K core matrices per step: +256 B
                    tc_mma(taddr, dAhi[buf] + o, dBhi[buf] + o, IDESC, (fresh && ks == 0) ? 0u : 1u);
                    tc_mma(taddr, dAlo[buf] + o, dBhi[buf] + o, IDESC, 1u);
                    tc_mma(taddr, dAhi[buf] + o, dBlo[buf] + o, IDESC, 1u);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(&bars[buf])) : "memory");
            }
            __syncwarp();
        }
    } else {
        // ------------------------------------------------ builder / drain warps
        auto prefetch = [&](int64_t c0, int buf) {
            const int cn = (int)min((int64_t)TC_PC, n_end - c0);
            float* rw = raw + buf * TC_PC * TC_RAWLD;
            for (int e = tid; e < TC_PC * (TC_D / 4); e += TC_THREADS) {
                const int p = e / (TC_D / 4), q4 = e - p * (TC_D / 4);
                float* dst = rw + p * TC_RAWLD + 4 * q4;
                if (p < cn) {
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tc_smem_u32(dst)), "l"(x + (c0 + p) * TC_D + 4 * q4) : "memory");
                } else {
                    *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            if (tid < TC_PC * 2) {
                const int p = tid >> 1, kk = tid & 1;
                float* dst = wraw + buf * TC_PC * 2 + tid;
                if (p < cn) {
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(tc_smem_u32(dst)), "l"(r + (c0 + p) * K + k0 + kk) : "memory");
                } else {
                    *dst = r_is_log ? -CUDART_INF_F : 0.f;
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        // this thread's items: feature j, point quads q0 + (TC_THREADS / 64) * h
        const int j = tid & 63, q0 = tid >> 6;
        const int toff = (j >> 3) * 256 + (j & 7) * 4;      // + q * 32: float offset of (row j, points 4q..4q+3) in a 64-row tile
        float sx[2] = {0.f, 0.f}, sw[2] = {0.f, 0.f};
        uint32_t uses[2] = {0, 0};
        bool ok = true;

        auto drain = [&](int lastbuf) {
            // every issued MMA is complete once the latest commit has arrived
            ok = ok && tc_wait(&bars[lastbuf], (uses[lastbuf] - 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (warp < 4) {
                uint32_t v[TC_D];
                const uint32_t a = taddr + ((uint32_t)(warp * 32) << 16);
#pragma unroll
                for (int c = 0; c < TC_D; c += 8) {
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                 : "=r"(v[c]), "=r"(v[c + 1]), "=r"(v[c + 2]), "=r"(v[c + 3]), "=r"(v[c + 4]), "=r"(v[c + 5]),
                                   "=r"(v[c + 6]), "=r"(v[c + 7])
                                 : "r"(a + c));
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                double* drow = dacc + (size_t)(warp * 32 + lane) * TC_DLD;       // row m = kk * 64 + i
#pragma unroll
                for (int c = 0; c < TC_D; ++c) drow[c] += (double)__uint_as_float(v[c]);
            }
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                atomicAdd(dsum + kk * (TC_D + 2) + j, (double)sx[kk]);
                if (j == 0) atomicAdd(dsum + kk * (TC_D + 2) + TC_D, (double)sw[kk]);
                sx[kk] = 0.f;
                sw[kk] = 0.f;
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            tc_bar_sync(1, TC_THREADS);
        };

        int lastbuf = 0;
        prefetch(n_begin, 0);
        for (int it = 0; it < nchunks; ++it) {
            const int buf = it & 1;
            const int64_t c0 = n_begin + (int64_t)it * TC_PC;
            if (it + 1 < nchunks) {
                prefetch(c0 + TC_PC, buf ^ 1);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            tc_bar_sync(1, TC_THREADS);
            if (tid < TC_PC * 2) {
                const float v = wraw[buf * TC_PC * 2 + tid];
                wsm[tid] = r_is_log ? expf(v) : v;
            }
            // the tensor core must be done with this tile set (chunk it-2) before it is overwritten
            if (uses[buf] > 0) ok = ok && tc_wait(&bars[buf], (uses[buf] - 1) & 1);
            tc_bar_sync(1, TC_THREADS);
            float* tb = tiles + buf * 6 * TC_TILE;
            const float* rw = raw + buf * TC_PC * TC_RAWLD;
#pragma unroll
            for (int h = 0; h < TC_ITEMS; ++h) {
                const int q = q0 + (TC_THREADS / 64) * h;
                float xv[4], hv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    xv[i] = rw[(4 * q + i) * TC_RAWLD + j];
                    hv[i] = tc_hi(xv[i]);
                }
                const int o = toff + q * 32;
                *reinterpret_cast<float4*>(tb + 4 * TC_TILE + o) = make_float4(hv[0], hv[1], hv[2], hv[3]);
                *reinterpret_cast<float4*>(tb + 5 * TC_TILE + o) = make_float4(xv[0] - hv[0], xv[1] - hv[1], xv[2] - hv[2], xv[3] - hv[3]);
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    float av[4], ah[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float w = wsm[(4 * q + i) * 2 + kk];
                        av[i] = w * xv[i];
                        ah[i] = tc_hi(av[i]);
                        sx[kk] += av[i];
                        if (j == 0) sw[kk] += w;
                    }
                    // component kk occupies rows 64 kk .. 64 kk + 63 of the stacked A tile = tile kk of the pair
                    *reinterpret_cast<float4*>(tb + (0 + kk) * TC_TILE + o) = make_float4(ah[0], ah[1], ah[2], ah[3]);
                    *reinterpret_cast<float4*>(tb + (2 + kk) * TC_TILE + o) = make_float4(av[0] - ah[0], av[1] - ah[1], av[2] - ah[2], av[3] - ah[3]);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy tile writes -> tensor core
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            tc_bar_arrive(2 + buf, TC_THREADS + 32);                       // hand the tile set to the issuer warp
            ++uses[buf];
            lastbuf = buf;
            if ((it + 1) % TC_FLUSH == 0 || it + 1 == nchunks) drain(lastbuf);
        }
        // an mbarrier wait that exhausted its spin budget means the tensor pipe never signalled: the accumulator is
        // incomplete — abort the launch (sticky CUDA error at the caller) rather than add wrong statistics
        if (!ok) __trap();

        // fp64 partials of this CTA -> global statistics (lower triangle mirrored: exactly symmetric)
        const int SL = stats_len(TC_D);
        for (int e = tid; e < 2 * TC_D * TC_D; e += TC_THREADS) {
            const int kk = e / (TC_D * TC_D), rem = e - kk * TC_D * TC_D, gi = rem / TC_D, gj = rem - gi * TC_D;
            if (gj > gi) continue;
            const double v = dacc[((size_t)kk * TC_D + gi) * TC_DLD + gj];
            double* out = stats + (size_t)(k0 + kk) * SL + 2 + TC_D;
            atomicAdd(out + gi * TC_D + gj, v);
            if (gi != gj) atomicAdd(out + gj * TC_D + gi, v);
        }
        if (tid < 2 * TC_D) {
            const int kk = tid / TC_D, jj = tid - kk * TC_D;
            atomicAdd(stats + (size_t)(k0 + kk) * SL + 2 + jj, dsum[kk * (TC_D + 2) + jj]);
        }
        if (tid < 2) {
            const double sw_tot = dsum[tid * (TC_D + 2) + TC_D];
            atomicAdd(stats + (size_t)(k0 + tid) * SL + 0, sw_tot);
            atomicAdd(stats + (size_t)(k0 + tid) * SL + 1, sw_tot);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(taddr) : "memory");
    ng_tail_run<float>(tail, K, TC_D, stats, gridDim.x * gridDim.y);
}

size_t suffstats_tc_smem_bytes() {
    return sizeof(float) * (2 * 6 * TC_TILE + 2 * TC_PC * TC_RAWLD + 2 * TC_PC * 2 + TC_PC * 2) +
           sizeof(double) * (2 * TC_D * TC_DLD + 2 * (TC_D + 2)) + 1024;
}

// returns VMP_OK, a CUDA error, or -100 when the shape does not qualify (caller falls back to the FP32 kernels)
int suffstats_tc(int64_t N, int K, int D, const float* x, const float* r, int r_is_log, double* stats, const NgTail& tail,
                 cudaStream_t st) {
    if (D != TC_D || (K & 1) || N < 4 * TC_PC) return -100;
    const size_t smem = suffstats_tc_smem_bytes();
    cudaError_t e = cudaFuncSetAttribute(suffstats_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int kpairs = K / 2;
    int nslices = (8 * 148 + kpairs - 1) / kpairs;
    const int64_t min_slice = 8 * TC_PC;
    if ((int64_t)nslices * min_slice > N) nslices = (int)((N + min_slice - 1) / min_slice);
    if (nslices < 1) nslices = 1;
    int64_t pps = (N + nslices - 1) / nslices;
    pps = ((pps + TC_PC - 1) / TC_PC) * TC_PC;
    nslices = (int)((N + pps - 1) / pps);
    suffstats_tc_kernel<<<dim3(kpairs, nslices), TC_THREADS + 32, smem, st>>>(N, K, pps, x, r, r_is_log, stats, tail);
    return launch_status();
}

}  // namespace vmp
