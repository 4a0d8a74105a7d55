// suffstats.cu — responsibility-weighted sufficient statistics, the fused natural-gradient (CVI) update and the
// standard-parameter M-step of the standalone mixtures.
//
// stats[k] = [ sum_n r_nk , sum_n w_nk , sum_n w_nk x_n (D) , sum_n w_nk x_n x_n^T (D*D) ]   (double, accumulated)
//   w = r (GMM / SVAE) or r*u (SMM).
// The kernel augments every point with a constant 1 (xt = [x, 1]) so that one register-tiled contraction
// sum_n w_nk xt xt^T yields the second moment, the first moment and the weight sum at once.  Each thread owns a
// 4x4 block of (i,j) for KT=2 components: 32 accumulators of type T flushed into 32 double accumulators every CH
// points (two-level summation keeps the fp32 rounding at the 1e-6 level for any N), and finally one double
// atomicAdd per output per CTA.
#include "common.cuh"
#include "ng_tail.cuh"


namespace vmp {

constexpr int SS_CH = 64;      // points per shared-memory chunk
constexpr int SS_KT = 2;       // components per thread

// (d0, d1) += a * (b0, b1): packed FFMA2 for float (two FMAs per issue slot on sm_100), plain FMAs for double
__device__ __forceinline__ void fma2_bcast(float& d0, float& d1, float a, float b0, float b1) {
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2,%2};\n\tmov.b64 rb, {%3,%4};\n\tmov.b64 rc, {%0,%1};\n\t"
        "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0,%1}, rc;\n\t}"
        : "+f"(d0), "+f"(d1)
        : "f"(a), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void fma2_bcast(double& d0, double& d1, double a, double b0, double b1) {
    d0 = fma(a, b0, d0);
    d1 = fma(a, b1, d1);
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
                 : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
                 : "memory");
}

// Thread t of a k-group owns the 4x4 block (bi, bj), bi >= bj, of the symmetric (D+1) x (D+1) matrix
// sum_n w_nk xt xt^T, xt = [x, 1] (lower block triangle only: the statistic is symmetric), for KT components.
// Chunks of SS_CH points stream through a double-buffered shared-memory stage filled by cp.async one chunk ahead;
// the augmented column [1, 0, ..] of the stage is written once and never overwritten.
template <typename T>
__global__ void __launch_bounds__(320)
suffstats_kernel(int64_t N, int K, int D, int D4, int nb, int G, int64_t pts_per_slice, const T* __restrict__ x,
                 const T* __restrict__ r, int r_is_log, const T* __restrict__ u_nk, double* __restrict__ stats,
                 const NgTail tail) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int KC = SS_KT * G;                       // components covered by this CTA
    T* xs0 = reinterpret_cast<T*>(smraw);           // [2][SS_CH][D4]   xt = [x, 1, 0 pad]
    T* ws0 = xs0 + 2 * (size_t)SS_CH * D4;          // [2][SS_CH][KC]   weights w (raw r / log r until converted)
    T* us0 = ws0 + 2 * (size_t)SS_CH * KC;          // [2][SS_CH][KC]   u (SMM) -> r after conversion
    const int k0 = blockIdx.x * KC;
    const int64_t n_begin = (int64_t)blockIdx.y * pts_per_slice;
    const int64_t n_end = min(N, n_begin + pts_per_slice);
    const int tid = threadIdx.x;
    const int nthreads = blockDim.x;
    const int nt = nb * (nb + 1) / 2;
    const int g = tid / nt, b = tid - g * nt;
    int bi = (int)((sqrtf(8.f * (float)b + 1.f) - 1.f) * 0.5f);
    while ((bi + 1) * (bi + 2) / 2 <= b) ++bi;
    while (bi * (bi + 1) / 2 > b) --bi;
    const int bj = b - bi * (bi + 1) / 2;
    const bool active = g < G && (k0 + g * SS_KT) < K;
    const bool vec16 = (D % 4 == 0) && sizeof(T) == 4;

    T acc[4][4][SS_KT];
    double dacc[4][4][SS_KT];
    T racc[SS_KT];
    double dracc[SS_KT];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int q = 0; q < SS_KT; ++q) { acc[i][j][q] = T(0); dacc[i][j][q] = 0.0; }
#pragma unroll
    for (int q = 0; q < SS_KT; ++q) { racc[q] = T(0); dracc[q] = 0.0; }

    // augmented / padding columns of both stage buffers
    for (int e = tid; e < 2 * SS_CH * (D4 - D); e += nthreads) {
        const int p = e / (D4 - D), i = D + e % (D4 - D);
        xs0[(size_t)p * D4 + i] = (i == D) ? T(1) : T(0);
    }

    auto prefetch = [&](int64_t c0, int buf) {
        const int cn = (int)min((int64_t)SS_CH, n_end - c0);
        T* xs = xs0 + (size_t)buf * SS_CH * D4;
        T* ws = ws0 + (size_t)buf * SS_CH * KC;
        T* us = us0 + (size_t)buf * SS_CH * KC;
        if (vec16) {
            const int per_row = D / 4;
            for (int e = tid; e < cn * per_row; e += nthreads) {
                const int p = e / per_row, q4 = e - p * per_row;
                cp_async16(xs + (size_t)p * D4 + 4 * q4, x + (c0 + p) * D + 4 * q4);
            }
        } else {
            for (int e = tid; e < cn * D; e += nthreads) {
                const int p = e / D, i = e - p * D;
                xs[(size_t)p * D4 + i] = x[(c0 + p) * D + i];
            }
        }
        for (int e = tid; e < cn * KC; e += nthreads) {
            const int p = e / KC, kk = e - p * KC;
            if (k0 + kk < K) {
                if (sizeof(T) == 4) {
                    cp_async4(ws + e, r + (c0 + p) * K + k0 + kk);
                    if (u_nk != nullptr) cp_async4(us + e, u_nk + (c0 + p) * K + k0 + kk);
                } else {
                    ws[e] = r[(c0 + p) * K + k0 + kk];
                    if (u_nk != nullptr) us[e] = u_nk[(c0 + p) * K + k0 + kk];
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    int it = 0;
    if (n_begin < n_end) prefetch(n_begin, 0);
    for (int64_t c0 = n_begin; c0 < n_end; c0 += SS_CH, ++it) {
        const int buf = it & 1;
        const int cn = (int)min((int64_t)SS_CH, n_end - c0);
        if (c0 + SS_CH < n_end) {
            prefetch(c0 + SS_CH, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        T* xs = xs0 + (size_t)buf * SS_CH * D4;
        T* ws = ws0 + (size_t)buf * SS_CH * KC;
        T* us = us0 + (size_t)buf * SS_CH * KC;
        // convert the weight tile in place: w = r (or exp(log r)) [* u], us <- r
        for (int e = tid; e < cn * KC; e += nthreads) {
            const int kk = e % KC;
            T rv = T(0), w = T(0);
            if (k0 + kk < K) {
                rv = ws[e];
                if (r_is_log) rv = t_exp(rv);
                w = u_nk != nullptr ? rv * us[e] : rv;
            }
            ws[e] = w;
            if (u_nk != nullptr) us[e] = rv;
        }
        __syncthreads();
        if (active) {
            for (int p = 0; p < cn; ++p) {
                const T* xp = xs + (size_t)p * D4;
                T xi[4], xj[4], w[SS_KT];
#pragma unroll
                for (int i = 0; i < 4; ++i) { xi[i] = xp[bi * 4 + i]; xj[i] = xp[bj * 4 + i]; }
#pragma unroll
                for (int q = 0; q < SS_KT; ++q) w[q] = ws[(size_t)p * KC + g * SS_KT + q];
#pragma unroll
                for (int q = 0; q < SS_KT; ++q)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const T wx = w[q] * xi[i];
                        fma2_bcast(acc[i][0][q], acc[i][1][q], wx, xj[0], xj[1]);
                        fma2_bcast(acc[i][2][q], acc[i][3][q], wx, xj[2], xj[3]);
                    }
                if (u_nk != nullptr && b == 0) {
#pragma unroll
                    for (int q = 0; q < SS_KT; ++q) racc[q] += us[(size_t)p * KC + g * SS_KT + q];
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int q = 0; q < SS_KT; ++q) { dacc[i][j][q] += (double)acc[i][j][q]; acc[i][j][q] = T(0); }
#pragma unroll
            for (int q = 0; q < SS_KT; ++q) { dracc[q] += (double)racc[q]; racc[q] = T(0); }
        }
        __syncthreads();      // everyone is done with this buffer before the next-but-one prefetch overwrites it
    }
    if (active) {
    const int SL = stats_len(D);
#pragma unroll
    for (int q = 0; q < SS_KT; ++q) {
        const int k = k0 + g * SS_KT + q;
        if (k >= K) continue;
        double* out = stats + (size_t)k * SL;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int gi = bi * 4 + i, gj = bj * 4 + j;
                const double v = dacc[i][j][q];
                if (gi < D && gj < D) {
                    if (gi >= gj) {                      // lower triangle only, mirrored: the statistic is exactly symmetric
                        atomicAdd(out + 2 + D + gi * D + gj, v);
                        if (gi != gj) atomicAdd(out + 2 + D + gj * D + gi, v);
                    }
                } else if (gi == D && gj < D) {
                    atomicAdd(out + 2 + gj, v);                                  // row of the augmented 1: sum w x
                } else if (gi == D && gj == D) {
                    atomicAdd(out + 1, v);
                    if (u_nk == nullptr) atomicAdd(out + 0, v);
                }
            }
        if (u_nk != nullptr && b == 0) atomicAdd(out + 0, dracc[q]);
    }
    }
    ng_tail_run<T>(tail, K, D, stats, gridDim.x * gridDim.y);
}

// ---- small latent dimension (D <= 8: the C1-C3 shapes): lane <-> component --------------------------------------
// A warp streams a run of points; lane l accumulates the (D+1)(D+2)/2 lower-triangular entries of w_nk xt xt^T for
// component k = 32*blockIdx.y + l in registers (x is a warp-uniform broadcast load, r[n, k..k+31] one coalesced
// 128-byte row).  fp32 partial sums cover at most SSM_RUN points, then go out as double atomics.
constexpr int SSM_RUN = 256;       // points per warp run (fp32 partial sums)
constexpr int SSM_WARPS = 8;

template <typename T, int D>
__global__ void __launch_bounds__(SSM_WARPS * 32)
suffstats_small_kernel(int64_t N, int K, const T* __restrict__ x, const T* __restrict__ r, int r_is_log,
                       const T* __restrict__ u_nk, double* __restrict__ stats, int64_t pts_per_warp, const NgTail tail) {
    constexpr int NA = (D + 1) * (D + 2) / 2;
    __shared__ double red[NA + 1][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp = (int64_t)blockIdx.x * SSM_WARPS + wib;
    const int k = blockIdx.y * 32 + lane;
    const int64_t n0 = min(N, warp * pts_per_warp), n1 = min(N, n0 + pts_per_warp);
    T acc[NA];
#pragma unroll
    for (int e = 0; e < NA; ++e) acc[e] = T(0);
    T racc = T(0);
    const bool kin = k < K;
    const T* rp = r + (kin ? k : 0);
    const T* up = u_nk != nullptr ? u_nk + (kin ? k : 0) : nullptr;
#pragma unroll 4
    for (int64_t n = n0; n < n1; ++n) {
        T xt[D + 1];
#pragma unroll
        for (int i = 0; i < D; ++i) xt[i] = x[n * D + i];
        xt[D] = T(1);
        T rv = rp[n * K];
        if (r_is_log) rv = t_exp(rv);
        if (!kin) rv = T(0);
        const T w = up != nullptr ? rv * up[n * K] : rv;
        racc += rv;
#pragma unroll
        for (int i = 0; i <= D; ++i) {
            const T wx = w * xt[i];
#pragma unroll
            for (int j = 0; j + 1 <= i; j += 2)
                fma2_bcast(acc[i * (i + 1) / 2 + j], acc[i * (i + 1) / 2 + j + 1], wx, xt[j], xt[j + 1]);
            if ((i & 1) == 0) acc[i * (i + 1) / 2 + i] = fma(wx, xt[i], acc[i * (i + 1) / 2 + i]);
        }
    }
    // CTA reduction in double (lane <-> component is the same in every warp), then one atomic per entry per CTA
    for (int wv = 0; wv < SSM_WARPS; ++wv) {
        if (wib == wv) {
#pragma unroll
            for (int e = 0; e < NA; ++e) red[e][lane] = (wv == 0 ? 0.0 : red[e][lane]) + (double)acc[e];
            red[NA][lane] = (wv == 0 ? 0.0 : red[NA][lane]) + (double)racc;
        }
        __syncthreads();
    }
    const int SL = stats_len(D);
    for (int t = threadIdx.x; t < (NA + 1) * 32; t += blockDim.x) {
        const int e = t >> 5, l = t & 31;
        const int kk = blockIdx.y * 32 + l;
        if (kk >= K) continue;
        double* out = stats + (size_t)kk * SL;
        const double v = red[e][l];
        if (e == NA) { atomicAdd(out + 0, v); continue; }
        int i = 0;
        while ((i + 1) * (i + 2) / 2 <= e) ++i;
        const int j = e - i * (i + 1) / 2;
        if (i < D) {
            atomicAdd(out + 2 + D + i * D + j, v);
            if (i != j) atomicAdd(out + 2 + D + j * D + i, v);
        } else if (j < D) {
            atomicAdd(out + 2 + j, v);
        } else {
            atomicAdd(out + 1, v);
        }
    }
    ng_tail_run<T>(tail, K, D, stats, gridDim.x * gridDim.y);
}

template <typename T, int D>
static int launch_suffstats_small(int64_t N, int K, const T* x, const T* r, int r_is_log, const T* u_nk, double* stats,
                                  const NgTail& tail, cudaStream_t st) {
    const int kblocks = (K + 31) / 32;
    int64_t ppw = SSM_RUN;
    const int64_t want_ctas = 148 * 2;
    if ((N + ppw * SSM_WARPS - 1) / (ppw * SSM_WARPS) < want_ctas)
        ppw = max((int64_t)8, (N + want_ctas * SSM_WARPS - 1) / (want_ctas * SSM_WARPS));
    const int64_t grid = (N + ppw * SSM_WARPS - 1) / (ppw * SSM_WARPS);
    if (grid > 0x7fffffffLL) return VMP_E_BADARG;
    suffstats_small_kernel<T, D><<<dim3((unsigned)grid, kblocks), SSM_WARPS * 32, 0, st>>>(N, K, x, r, r_is_log, u_nk,
                                                                                          stats, ppw, tail);
    return launch_status();
}

// tensor-core contraction (suffstats_tc.cu): fp32, D = 64, even K, GMM weights, N >= 128
int suffstats_tc(int64_t N, int K, int D, const float* x, const float* r, int r_is_log, double* stats, const NgTail& tail,
                 cudaStream_t st);
template <typename T>
static int suffstats_tc_dispatch(int64_t, int, int, const T*, const T*, int, const T*, double*, const NgTail&, cudaStream_t) { return -100; }
template <>
int suffstats_tc_dispatch<float>(int64_t N, int K, int D, const float* x, const float* r, int r_is_log, const float* u_nk,
                                 double* stats, const NgTail& tail, cudaStream_t st) {
    if (u_nk != nullptr) return -100;
    return suffstats_tc(N, K, D, x, r, r_is_log, stats, tail, st);
}

// warp-level tensor-core contraction (suffstats_mma.cu): fp32, D in {16, 32}, K % 4 == 0, GMM or SMM weights
int suffstats_mma(int64_t N, int K, int D, const float* x, const float* r, int r_is_log, const float* u, double* stats,
                  const NgTail& tail, cudaStream_t st);
template <typename T>
static int suffstats_mma_dispatch(int64_t, int, int, const T*, const T*, int, const T*, double*, const NgTail&, cudaStream_t) { return -100; }
template <>
int suffstats_mma_dispatch<float>(int64_t N, int K, int D, const float* x, const float* r, int r_is_log, const float* u_nk,
                                  double* stats, const NgTail& tail, cudaStream_t st) {
    return suffstats_mma(N, K, D, x, r, r_is_log, u_nk, stats, tail, st);
}

// fp32, D <= 8, K <= 32, plain responsibilities: the lane <-> component kernel of mixture_sweep.cu
int sweep_stats_f32(int64_t N, int K, int D, const float* x, const float* r, const float* u, double* stats, const NgTail& tail,
                    cudaStream_t st);
template <typename T>
static int sweep_stats_dispatch(int64_t, int, int, const T*, const T*, int, const T*, double*, const NgTail&, cudaStream_t) { return -100; }
template <>
int sweep_stats_dispatch<float>(int64_t N, int K, int D, const float* x, const float* r, int r_is_log, const float* u_nk,
                                double* stats, const NgTail& tail, cudaStream_t st) {
    if (r_is_log || D > 8 || K > 32) return -100;
    return sweep_stats_f32(N, K, D, x, r, u_nk, stats, tail, st);
}

template <typename T>
int suffstats(int64_t N, int K, int D, const T* x, const T* r, int r_is_log, const T* u_nk, double* stats,
              const NgTail& tail, void* stream) {
    if (N < 0 || K <= 0) return VMP_E_BADARG;
    if (D < 1 || D > VMP_MAX_D) return VMP_E_BADDIM;
    if (N == 0) return VMP_OK;
    if (!x || !r || !stats) return VMP_E_BADARG;
    if (int rc = sweep_stats_dispatch(N, K, D, x, r, r_is_log, u_nk, stats, tail, (cudaStream_t)stream); rc != -100) return rc;
    if (int rc = suffstats_tc_dispatch(N, K, D, x, r, r_is_log, u_nk, stats, tail, (cudaStream_t)stream); rc != -100) return rc;
    if (int rc = suffstats_mma_dispatch(N, K, D, x, r, r_is_log, u_nk, stats, tail, (cudaStream_t)stream); rc != -100) return rc;
#define VMP_SSM(DD) \
    case DD: return launch_suffstats_small<T, DD>(N, K, x, r, r_is_log, u_nk, stats, tail, (cudaStream_t)stream)
    switch (D) {
        VMP_SSM(1); VMP_SSM(2); VMP_SSM(3); VMP_SSM(4); VMP_SSM(5); VMP_SSM(6); VMP_SSM(7); VMP_SSM(8);
        default: break;
    }
#undef VMP_SSM
    const int D4 = ((D + 1 + 3) / 4) * 4;
    const int nb = D4 / 4;
    const int nt = nb * (nb + 1) / 2;
    int G = 256 / nt;
    if (G < 1) G = 1;
    const int kgroups = (K + SS_KT - 1) / SS_KT;
    if (G > kgroups) G = kgroups;
    const int threads = ((nt * G + 31) / 32) * 32;
    const int KC = SS_KT * G;
    const int ktiles = (K + KC - 1) / KC;
    int nslices = (6 * 148 + ktiles - 1) / ktiles;
    const int64_t min_slice = 4 * SS_CH;
    if ((int64_t)nslices * min_slice > N) nslices = (int)((N + min_slice - 1) / min_slice);
    if (nslices < 1) nslices = 1;
    int64_t pps = (N + nslices - 1) / nslices;
    pps = ((pps + SS_CH - 1) / SS_CH) * SS_CH;
    nslices = (int)((N + pps - 1) / pps);
    const size_t smem = sizeof(T) * 2 * ((size_t)SS_CH * D4 + 2 * (size_t)SS_CH * KC);
    auto kern = suffstats_kernel<T>;
    if (smem + 256 > 48 * 1024) {          // the kernel also has a few bytes of static shared memory (ticket of the fused tail)
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    kern<<<dim3(ktiles, nslices), threads, smem, (cudaStream_t)stream>>>(N, K, D, D4, nb, G, pps, x, r, r_is_log, u_nk,
                                                                        stats, tail);
    return launch_status();
}

// theta* = prior + [N_k, sum r x x^T, sum r x, N_k, N_k + 1]   (svae.m_step in natural parameters)
// theta <- (1 - rho) theta + rho theta*                         (svae.update_gmm_params)
template <typename T>
__global__ void ng_update_kernel(int K, int D, const double* __restrict__ stats, double rho_host,
                                 const double* __restrict__ rho_dev, int only_alpha,
                                 const T* __restrict__ p_alpha, const T* __restrict__ p_A, const T* __restrict__ p_b,
                                 const T* __restrict__ p_beta, const T* __restrict__ p_vhat, T* alpha, T* A, T* b,
                                 T* beta, T* v_hat, T* s_alpha, T* s_A, T* s_b, T* s_beta, T* s_vhat) {
    const int k = blockIdx.x;
    const double rho = rho_dev != nullptr ? *rho_dev : rho_host;     // device-resident step size: CUDA-graph replays
    const double* st = stats + (size_t)k * stats_len(D);
    const double Nk = st[0];
    if (threadIdx.x == 0) {
        const double a_star = (double)p_alpha[k] + Nk;
        if (s_alpha) s_alpha[k] = (T)a_star;
        alpha[k] = (T)((1.0 - rho) * (double)alpha[k] + rho * a_star);
        if (!only_alpha) {
            const double be = (double)p_beta[k] + Nk, vh = (double)p_vhat[k] + Nk + 1.0;
            if (s_beta) s_beta[k] = (T)be;
            if (s_vhat) s_vhat[k] = (T)vh;
            beta[k] = (T)((1.0 - rho) * (double)beta[k] + rho * be);
            v_hat[k] = (T)((1.0 - rho) * (double)v_hat[k] + rho * vh);
        }
    }
    if (only_alpha) return;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        const size_t o = (size_t)k * D + i;
        const double v = (double)p_b[o] + st[2 + i];
        if (s_b) s_b[o] = (T)v;
        b[o] = (T)((1.0 - rho) * (double)b[o] + rho * v);
    }
    for (int e = threadIdx.x; e < D * D; e += blockDim.x) {
        const size_t o = (size_t)k * D * D + e;
        const double v = (double)p_A[o] + st[2 + D + e];
        if (s_A) s_A[o] = (T)v;
        A[o] = (T)((1.0 - rho) * (double)A[o] + rho * v);
    }
}

template <typename T>
int ng_update(int K, int D, const double* stats, double rho, const double* rho_dev, int only_alpha, const T* p_alpha, const T* p_A,
              const T* p_b, const T* p_beta, const T* p_vhat, T* alpha, T* A, T* b, T* beta, T* v_hat, T* s_alpha,
              T* s_A, T* s_b, T* s_beta, T* s_vhat, void* stream) {
    if (K <= 0 || !stats || !p_alpha || !alpha) return VMP_E_BADARG;
    if (!only_alpha && (!p_A || !p_b || !p_beta || !p_vhat || !A || !b || !beta || !v_hat)) return VMP_E_BADARG;
    if (D < 1 || D > VMP_MAX_D) return VMP_E_BADDIM;
    ng_update_kernel<T><<<K, 256, 0, (cudaStream_t)stream>>>(K, D, stats, rho, rho_dev, only_alpha, p_alpha, p_A, p_b, p_beta,
                                                             p_vhat, alpha, A, b, beta, v_hat, s_alpha, s_A, s_b,
                                                             s_beta, s_vhat);
    return launch_status();
}

// Standard-parameter M-step of the standalone mixtures from the additive statistics.
template <typename T>
__global__ void mixture_mstep_kernel(int K, int D, int is_smm, const double* __restrict__ stats,
                                     const T* __restrict__ alpha_0, const T* __restrict__ beta_0,
                                     const T* __restrict__ m_0, const T* __restrict__ C_0, const T* __restrict__ v_0,
                                     T* alpha_k, T* beta_k, T* m_k, T* C_k, T* v_k, T* x_k, T* S_k) {
    const int k = blockIdx.x;
    const double* st = stats + (size_t)k * stats_len(D);
    const double Nk = st[0], Wk = st[1];
    const double* s1 = st + 2;
    const double* s2 = st + 2 + D;
    // gmm.py:30-36 NaN guard (N_k == 0 -> unnormalised sums) / smm.py:35-40 eps
    const double den = is_smm ? Wk + 1e-20 : Wk;
    const bool raw = !is_smm && !(Wk != 0.0);
    const double b0 = (double)beta_0[k];
    const double bk = b0 + Wk;
    if (threadIdx.x == 0) {
        alpha_k[k] = (T)((double)alpha_0[k] + Nk);
        beta_k[k] = (T)bk;
        v_k[k] = (T)((double)v_0[k] + Nk + (is_smm ? 0.0 : 1.0));      // gmm.py:81 (+1) vs smm.py:76
    }
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        const double xk = raw ? s1[i] : s1[i] / den;
        x_k[(size_t)k * D + i] = (T)xk;
        m_k[(size_t)k * D + i] = (T)((b0 * (double)m_0[(size_t)k * D + i] + Wk * xk) / bk);
    }
    for (int e = threadIdx.x; e < D * D; e += blockDim.x) {
        const int i = e / D, j = e - i * D;
        const double xi = raw ? s1[i] : s1[i] / den, xj = raw ? s1[j] : s1[j] / den;
        // sum w (x - x_k)(x - x_k)^T expanded around the additive moments
        const double Sraw = s2[e] - xi * s1[j] - s1[i] * xj + Wk * xi * xj;
        const double S = raw ? Sraw : Sraw / den;
        S_k[(size_t)k * D * D + e] = (T)S;
        const double qi = xi - (double)m_0[(size_t)k * D + i], qj = xj - (double)m_0[(size_t)k * D + j];
        C_k[(size_t)k * D * D + e] = (T)((double)C_0[(size_t)k * D * D + e] + Wk * S + b0 * Wk / bk * qi * qj);
    }
}

template <typename T>
int mixture_mstep(int K, int D, int is_smm, const double* stats, const T* alpha_0, const T* beta_0, const T* m_0,
                  const T* C_0, const T* v_0, T* alpha_k, T* beta_k, T* m_k, T* C_k, T* v_k, T* x_k, T* S_k,
                  void* stream) {
    if (K <= 0 || !stats || !alpha_0 || !beta_0 || !m_0 || !C_0 || !v_0 || !alpha_k || !beta_k || !m_k || !C_k ||
        !v_k || !x_k || !S_k)
        return VMP_E_BADARG;
    if (D < 1 || D > VMP_MAX_D) return VMP_E_BADDIM;
    mixture_mstep_kernel<T><<<K, 256, 0, (cudaStream_t)stream>>>(K, D, is_smm, stats, alpha_0, beta_0, m_0, C_0, v_0,
                                                                 alpha_k, beta_k, m_k, C_k, v_k, x_k, S_k);
    return launch_status();
}

}  // namespace vmp

extern "C" {
int vmp_suffstats_f32(int64_t N, int K, int D, const float* x, const float* r, int r_is_log, const float* u_nk,
                      double* stats, void* stream) {
    return vmp::suffstats<float>(N, K, D, x, r, r_is_log, u_nk, stats, vmp::ng_tail_none(), stream);
}
int vmp_suffstats_f64(int64_t N, int K, int D, const double* x, const double* r, int r_is_log, const double* u_nk,
                      double* stats, void* stream) {
    return vmp::suffstats<double>(N, K, D, x, r, r_is_log, u_nk, stats, vmp::ng_tail_none(), stream);
}
static vmp::NgTail make_tail(unsigned int* counter, double rho, const double* rho_dev, int only_alpha, const void* p_alpha,
                             const void* p_A, const void* p_b, const void* p_beta, const void* p_vhat, void* alpha, void* A,
                             void* b, void* beta, void* v_hat) {
    vmp::NgTail t;
    t.counter = counter; t.rho = rho; t.rho_dev = rho_dev; t.only_alpha = only_alpha;
    t.p_alpha = p_alpha; t.p_A = p_A; t.p_b = p_b; t.p_beta = p_beta; t.p_vhat = p_vhat;
    t.alpha = alpha; t.A = A; t.b = b; t.beta = beta; t.v_hat = v_hat;
    return t;
}
static bool tail_ok(const vmp::NgTail& t) {
    if (!t.counter || !t.p_alpha || !t.alpha) return false;
    return t.only_alpha || (t.p_A && t.p_b && t.p_beta && t.p_vhat && t.A && t.b && t.beta && t.v_hat);
}
int vmp_suffstats_update_f32(int64_t N, int K, int D, const float* x, const float* r, int r_is_log, const float* u_nk,
                             double* stats, unsigned int* counter, double rho, const double* rho_dev, int only_alpha,
                             const float* p_alpha, const float* p_A, const float* p_b, const float* p_beta,
                             const float* p_vhat, float* alpha, float* A, float* b, float* beta, float* v_hat, void* stream) {
    const vmp::NgTail t = make_tail(counter, rho, rho_dev, only_alpha, p_alpha, p_A, p_b, p_beta, p_vhat, alpha, A, b, beta, v_hat);
    if (N <= 0 || !tail_ok(t)) return VMP_E_BADARG;
    return vmp::suffstats<float>(N, K, D, x, r, r_is_log, u_nk, stats, t, stream);
}
int vmp_suffstats_update_f64(int64_t N, int K, int D, const double* x, const double* r, int r_is_log, const double* u_nk,
                             double* stats, unsigned int* counter, double rho, const double* rho_dev, int only_alpha,
                             const double* p_alpha, const double* p_A, const double* p_b, const double* p_beta,
                             const double* p_vhat, double* alpha, double* A, double* b, double* beta, double* v_hat,
                             void* stream) {
    const vmp::NgTail t = make_tail(counter, rho, rho_dev, only_alpha, p_alpha, p_A, p_b, p_beta, p_vhat, alpha, A, b, beta, v_hat);
    if (N <= 0 || !tail_ok(t)) return VMP_E_BADARG;
    return vmp::suffstats<double>(N, K, D, x, r, r_is_log, u_nk, stats, t, stream);
}
int vmp_ng_update_f32(int K, int D, const double* stats, double rho, const double* rho_dev, int only_alpha,
                      const float* p_alpha,
                      const float* p_A, const float* p_b, const float* p_beta, const float* p_vhat, float* alpha,
                      float* A, float* b, float* beta, float* v_hat, float* s_alpha, float* s_A, float* s_b,
                      float* s_beta, float* s_vhat, void* stream) {
    return vmp::ng_update<float>(K, D, stats, rho, rho_dev, only_alpha, p_alpha, p_A, p_b, p_beta, p_vhat, alpha, A, b, beta,
                                 v_hat, s_alpha, s_A, s_b, s_beta, s_vhat, stream);
}
int vmp_ng_update_f64(int K, int D, const double* stats, double rho, const double* rho_dev, int only_alpha,
                      const double* p_alpha,
                      const double* p_A, const double* p_b, const double* p_beta, const double* p_vhat, double* alpha,
                      double* A, double* b, double* beta, double* v_hat, double* s_alpha, double* s_A, double* s_b,
                      double* s_beta, double* s_vhat, void* stream) {
    return vmp::ng_update<double>(K, D, stats, rho, rho_dev, only_alpha, p_alpha, p_A, p_b, p_beta, p_vhat, alpha, A, b, beta,
                                  v_hat, s_alpha, s_A, s_b, s_beta, s_vhat, stream);
}
int vmp_mixture_mstep_f32(int K, int D, int is_smm, const double* stats, const float* alpha_0, const float* beta_0,
                          const float* m_0, const float* C_0, const float* v_0, float* alpha_k, float* beta_k,
                          float* m_k, float* C_k, float* v_k, float* x_k, float* S_k, void* stream) {
    return vmp::mixture_mstep<float>(K, D, is_smm, stats, alpha_0, beta_0, m_0, C_0, v_0, alpha_k, beta_k, m_k, C_k,
                                     v_k, x_k, S_k, stream);
}
int vmp_mixture_mstep_f64(int K, int D, int is_smm, const double* stats, const double* alpha_0, const double* beta_0,
                          const double* m_0, const double* C_0, const double* v_0, double* alpha_k, double* beta_k,
                          double* m_k, double* C_k, double* v_k, double* x_k, double* S_k, void* stream) {
    return vmp::mixture_mstep<double>(K, D, is_smm, stats, alpha_0, beta_0, m_0, C_0, v_0, alpha_k, beta_k, m_k, C_k,
                                      v_k, x_k, S_k, stream);
}
}
