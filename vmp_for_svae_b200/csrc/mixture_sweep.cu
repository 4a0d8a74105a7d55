// mixture_sweep.cu — whole VB-EM sweeps of the standalone mixtures (gmm.inference gmm.py:230-269, smm.inference
// smm.py:199-245) and the driver loops around them (gmm.py:377-379, smm.py `for i in range(nb_iters): sess.run(update)`).
//
// One sweep is  m_step(r[,u]) -> P = inv(C) -> e_step -> (r[,u]) .  Kernels:
//   S  sweep_stats_kernel   : statistics [sum r, sum w, sum w x, sum w x x^T] of a GIVEN state (r, u)       (first M-step)
//   P  sweep_prepare_kernel : K CTAs, double: M-step in standard parameters (gmm.py:25-81 / smm.py:25-85), Cholesky of C_k,
//                             P_k = C_k^-1, the per-component E-step constants (gmm.py:117-138 / smm.py:100-128) and one packed
//                             fp32 record per component; also zeroes the statistics buffer of the NEXT sweep
//   E  sweep_estep_kernel   : lane <-> component (K <= 32), the component's record lives in REGISTERS for the whole kernel;
//                             a warp streams points: x row broadcast from a per-warp staging tile, 44 FMAs for the expected
//                             Mahalanobis distance, softmax across the lanes with shuffles, r / u written as coalesced 128-byte
//                             rows — and, fused, the statistics of the NEW (r, u) for the next sweep's M-step, so that inside
//                             a multi-sweep run r and u never travel through HBM (only the last sweep writes them).
// Algorithmic HBM bytes of one sweep with state in / state out: 4 (D + 4K) per point (read x, r, u; write r, u);
// inside vmp_mixture_fit every sweep but the last reads only x.
// fp32, D <= 8, K <= 32 run these kernels; everything else (fp64, larger D or K, missing-data masks) goes through the
// general kernels of suffstats.cu / mixtures.cu from the same entry point.
#include <type_traits>

#include "block_linalg.cuh"
#include "common.cuh"
#include "ng_tail.cuh"

namespace vmp {

constexpr int SW_WARPS = 8;                 // warps per CTA of the two-CTA-per-SM kernels
constexpr int SW_WARPS_FUSED = 8;           // e-step + statistics: ~200 registers, one CTA per SM (8 independent chains per warp)
constexpr int SW_RUN = 256;                 // points per fp32 partial-sum run
constexpr float SW_LOG2E = 1.4426950408889634f;

__host__ __device__ constexpr int sw_np(int D) { return D * (D + 1) / 2; }
__host__ __device__ constexpr int sw_rs(int D) { return ((sw_np(D) + D + 4 + 3) / 4) * 4; }
__host__ __device__ constexpr int sw_na(int D) { return (D + 1) * (D + 2) / 2; }

// packed FP32x2 FMA (sm_100 FFMA2): (d0, d1) += a * (b0, b1)
__device__ __forceinline__ void sw_ffma2_bcast(float& d0, float& d1, float a, float b0, float b1) {
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2,%2};\n\tmov.b64 rb, {%3,%4};\n\tmov.b64 rc, {%0,%1};\n\t"
        "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0,%1}, rc;\n\t}"
        : "+f"(d0), "+f"(d1)
        : "f"(a), "f"(b0), "f"(b1));
}

// (d0, d1) += (a0, a1) * (b0, b1)
__device__ __forceinline__ void sw_ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2,%3};\n\tmov.b64 rb, {%4,%5};\n\tmov.b64 rc, {%0,%1};\n\t"
        "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0,%1}, rc;\n\t}"
        : "+f"(d0), "+f"(d1)
        : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
// (d0, d1) = (a0, a1) + (b, b)
__device__ __forceinline__ void sw_fadd2_bcast(float& d0, float& d1, float a0, float a1, float b) {
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2,%3};\n\tmov.b64 rb, {%4,%4};\n\t"
        "add.rn.f32x2 rc, ra, rb;\n\tmov.b64 {%0,%1}, rc;\n\t}"
        : "=f"(d0), "=f"(d1)
        : "f"(a0), "f"(a1), "f"(b));
}
// warp-wide float max in ONE instruction: order-preserving map to int32 + REDUX.MAX.S32
__device__ __forceinline__ float sw_warp_max(float v) {
    int k = __float_as_int(v);
    k = k >= 0 ? k : k ^ 0x7fffffff;
    k = __reduce_max_sync(0xffffffffu, k);
    k = k >= 0 ? k : k ^ 0x7fffffff;
    return __int_as_float(k);
}
__device__ __forceinline__ float sw_ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void sw_cp_async16(float* dst, const float* src, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    const int sz = valid ? 16 : 0;                                   // src-size 0: zero-fill, nothing is read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void sw_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void sw_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---------------------------------------------------------------------------------------------------- P: prologue
// stats[k] = [N_k, W_k, sum w x (D), sum w x x^T (D*D)] (double).  Outputs in T; rec (fp32 packed records) optional.
template <typename T>
__global__ void __launch_bounds__(128)
sweep_prepare_kernel(int K, int D, int is_smm, const double* __restrict__ stats, double* __restrict__ stats_next,
                     const T* __restrict__ alpha_0, const T* __restrict__ beta_0, const T* __restrict__ m_0,
                     const T* __restrict__ C_0, const T* __restrict__ v_0, const T* __restrict__ kappa_k,
                     T* __restrict__ alpha_k, T* __restrict__ beta_k, T* __restrict__ m_k, T* __restrict__ C_k,
                     T* __restrict__ v_k, T* __restrict__ x_k, T* __restrict__ S_k, T* __restrict__ pi,
                     T* __restrict__ P_k, T* __restrict__ cst, float* __restrict__ rec) {
    extern __shared__ double sm[];
    const int ld = D + 1;
    double* C = sm;                 // [D][ld]  C_k -> its Cholesky factor
    double* W = C + D * ld;         // [D][ld]  Lc^-1
    double* xk = W + D * ld;        // [D]
    double* mk = xk + D;            // [D]
    double* red = mk + D;           // [32]
    const int k = blockIdx.x, SL = stats_len(D);
    const double* st = stats + (size_t)k * SL;
    const double Nk = st[0], Wk = st[1];
    const double* s1 = st + 2;
    const double* s2 = st + 2 + D;
    // gmm.py:30-36,42-46 NaN guard (N_k == 0 -> unnormalised sums) / smm.py:35-50 eps
    const double den = is_smm ? Wk + 1e-20 : Wk;
    const bool raw = !is_smm && !(Wk != 0.0);
    const double b0 = (double)beta_0[k], bk = b0 + Wk;
    const double vk = (double)v_0[k] + Nk + (is_smm ? 0.0 : 1.0);             // gmm.py:81 (+1) vs smm.py:76
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        const double xi = raw ? s1[i] : s1[i] / den;
        xk[i] = xi;
        mk[i] = (b0 * (double)m_0[(size_t)k * D + i] + Wk * xi) / bk;
        x_k[(size_t)k * D + i] = (T)xi;
        m_k[(size_t)k * D + i] = (T)mk[i];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < D * D; e += blockDim.x) {
        const int i = e / D, j = e - i * D;
        const double Sraw = s2[e] - xk[i] * s1[j] - s1[i] * xk[j] + Wk * xk[i] * xk[j];
        const double S = raw ? Sraw : Sraw / den;
        S_k[(size_t)k * D * D + e] = (T)S;
        const double qi = xk[i] - (double)m_0[(size_t)k * D + i], qj = xk[j] - (double)m_0[(size_t)k * D + j];
        const double c = (double)C_0[(size_t)k * D * D + e] + Wk * S + b0 * Wk / bk * qi * qj;
        C_k[(size_t)k * D * D + e] = (T)c;
        C[i * ld + j] = c;
    }
    // sum_j alpha_j of the UPDATED Dirichlet (gmm.py:134-138)
    double sa = 0.0;
    for (int j = threadIdx.x; j < K; j += blockDim.x) sa += (double)alpha_0[j] + stats[(size_t)j * SL];
    sa = block_sum(sa, red);
    __shared__ double sa_s;
    if (threadIdx.x == 0) sa_s = sa;
    __syncthreads();
    sa = sa_s;
    chol_lower_block(C, D, ld);
    tri_inverse_block(C, W, D, ld);
    const double ak = (double)alpha_0[k] + Nk;
    // the special functions are spread over the threads (each is a few hundred cycles of serial double arithmetic):
    // thread i < D: log L_ii and psi((v [+1] + i) / 2); thread D .. D+3: psi(alpha_k), psi(sum alpha), the two lgammas
    const double kap = is_smm ? (double)kappa_k[k] : 0.0;
    double part = 0.0;
    {
        const int t = threadIdx.x;
        if (t < D) {
            const double psi = digamma_pos(0.5 * (vk + (is_smm ? 0.0 : 1.0) + t));            // gmm.py:128-129 / smm.py:108
            part = 0.5 * psi + (is_smm ? -log(C[t * ld + t]) : 0.0);                           // 1/2 logdet P = -sum log Lc_ii
            xk[t] = log(C[t * ld + t]);                                                        // (xk is free again) for the GMM guard
        } else if (t == D) {
            part = digamma_pos(ak);
        } else if (t == D + 1) {
            part = -digamma_pos(sa);
        } else if (t == D + 2 && is_smm) {
            part = lgamma(0.5 * (D + kap));
        } else if (t == D + 3 && is_smm) {
            part = -lgamma(0.5 * kap);
        }
    }
    const double tot = block_sum(part, red);
    __shared__ double tot_s, elogpi_s;
    if (threadIdx.x == D) elogpi_s = part;
    __syncthreads();
    if (threadIdx.x == D + 1) elogpi_s += part;
    if (threadIdx.x == 0) tot_s = tot;
    __syncthreads();
    const double elogpi = elogpi_s;
    double logdetC = 0.0;
    for (int i = 0; i < D; ++i) logdetC += 2.0 * xk[i];
    const double logdetP = -logdetC;
    double c, hk, ek = 0.0, Dk = 0.0;
    const double dbeta = (double)D / bk;
    if (!is_smm) {
        const double ld_guard = (logdetP > -46.051701859880914) ? logdetP : 0.0;              // det > 1e-20 (gmm.py:120-121)
        c = tot_s + 0.5 * (D * VMP_LOG_2 + ld_guard);                                         // E log pi + 1/2 E log|Lambda|
        hk = 0.5;
    } else {
        // smm.py:119-124: lgamma((D+kap)/2) - lgamma(kap/2) - D/2 log(kap pi) + E log pi + 1/2 E log|Lambda| + log kap
        c = tot_s + 0.5 * D * VMP_LOG_2 - 0.5 * D * log(kap * 3.14159265358979323846) + log(kap);
        hk = 0.5 * (D + kap);
        ek = dbeta + kap;
        Dk = D + kap;
    }
    if (threadIdx.x == 0) {
        alpha_k[k] = (T)ak;
        beta_k[k] = (T)bk;
        v_k[k] = (T)vk;
        pi[k] = (T)exp(elogpi);
        if (cst) cst[k] = (T)c;
    }
    // P = W^T W ; record: tri (v P, strictly-lower entries doubled) | -m | log2(e) (c - hk D/beta) | log2(e) hk |
    // D/beta + kappa | D + kappa     (the e-step evaluates the soft-max in base 2: one FFMA + one MUFU.EX2 per pair)
    const int NP = sw_np(D), RS = sw_rs(D);
    for (int e = threadIdx.x; e < D * D; e += blockDim.x) {
        const int i = e / D, j = e - i * D;
        const int m = i > j ? i : j;
        double s = 0.0;
        for (int q = m; q < D; ++q) s += W[q * ld + i] * W[q * ld + j];
        if (P_k) P_k[(size_t)k * D * D + e] = (T)s;
        if (rec && j <= i) rec[(size_t)k * RS + i * (i + 1) / 2 + j] = (float)(vk * (i == j ? s : 2.0 * s));
    }
    if (rec) {
        for (int i = threadIdx.x; i < D; i += blockDim.x) rec[(size_t)k * RS + NP + i] = (float)(-mk[i]);
        if (threadIdx.x == 0) {
            float* o = rec + (size_t)k * RS + NP + D;
            o[0] = (float)(1.4426950408889634074 * (c - hk * dbeta));
            o[1] = (float)(1.4426950408889634074 * hk);
            o[2] = (float)ek;
            o[3] = (float)Dk;
            for (int e = NP + D + 4; e < RS; ++e) rec[(size_t)k * RS + e] = 0.f;
        }
    }
    if (stats_next) for (int e = threadIdx.x; e < SL; e += blockDim.x) stats_next[(size_t)k * SL + e] = 0.0;
}

// ---------------------------------------------------------------------------------------------------- shared device code
// accumulate w * xt xt^T (lower triangle of the augmented xt = [x, 1]) into acc, packed pairs
template <int D>
__device__ __forceinline__ void sw_accumulate(float (&acc)[sw_na(D)], const float (&xt)[D + 1], float w) {
#pragma unroll
    for (int i = 0; i <= D; ++i) {
        const float wx = w * xt[i];
#pragma unroll
        for (int j = 0; j + 1 <= i; j += 2)
            sw_ffma2_bcast(acc[i * (i + 1) / 2 + j], acc[i * (i + 1) / 2 + j + 1], wx, xt[j], xt[j + 1]);
        if ((i & 1) == 0) acc[i * (i + 1) / 2 + i] = fmaf(wx, xt[i], acc[i * (i + 1) / 2 + i]);
    }
}

template <int D>
__device__ __forceinline__ void sw_flush(float (&acc)[sw_na(D)], float& racc, double (*red)[32], int lane) {
    constexpr int NA = sw_na(D);
#pragma unroll
    for (int e = 0; e < NA; ++e) {
        atomicAdd(&red[e][lane], (double)acc[e]);
        acc[e] = 0.f;
    }
    atomicAdd(&red[NA][lane], (double)racc);
    racc = 0.f;
}

template <int D>
__device__ __forceinline__ void sw_store_stats(double (*red)[32], int K, bool weighted, double* __restrict__ stats) {
    constexpr int NA = sw_na(D);
    const int SL = stats_len(D);
    for (int t = threadIdx.x; t < (NA + 1) * 32; t += blockDim.x) {
        const int e = t >> 5, l = t & 31;
        if (l >= K) continue;
        double* out = stats + (size_t)l * SL;
        const double v = red[e][l];
        if (v == 0.0) continue;
        if (e == NA) { atomicAdd(out + 0, v); continue; }          // sum r
        int i = 0;
        while ((i + 1) * (i + 2) / 2 <= e) ++i;
        const int j = e - i * (i + 1) / 2;
        if (i < D) {
            atomicAdd(out + 2 + D + i * D + j, v);
            if (i != j) atomicAdd(out + 2 + D + j * D + i, v);
        } else if (j < D) {
            atomicAdd(out + 2 + j, v);
        } else {
            atomicAdd(out + 1, v);                                   // sum w
        }
    }
    (void)weighted;
}

// stage the x rows of up to 32 consecutive points (contiguous 32*D floats) into the warp's tile
template <int D>
__device__ __forceinline__ void sw_load_x(const float* __restrict__ x, int64_t n0, int cnt, float (&pre)[D], int lane) {
#pragma unroll
    for (int j = 0; j < D; ++j) {
        const int e = lane + 32 * j;
        pre[j] = e < cnt * D ? x[n0 * D + e] : 0.f;
    }
}
template <int D>
__device__ __forceinline__ void sw_put_x(float* xs, const float (&pre)[D], int lane) {
#pragma unroll
    for (int j = 0; j < D; ++j) xs[lane + 32 * j] = pre[j];
}
template <int D>
__device__ __forceinline__ void sw_get_row(const float* xs, int p, float (&xt)[D + 1]) {
    if constexpr (D % 4 == 0) {
#pragma unroll
        for (int q = 0; q < D / 4; ++q) {
            const float4 v = reinterpret_cast<const float4*>(xs + p * D)[q];
            xt[4 * q] = v.x; xt[4 * q + 1] = v.y; xt[4 * q + 2] = v.z; xt[4 * q + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < D; ++i) xt[i] = xs[p * D + i];
    }
    xt[D] = 1.f;
}

// pair-interleaved tile for the packed e-step: element (p, i) at (p / 2) * 2D + 2 i + (p & 1), so that one 64-bit read yields
// (x_{2q}[i], x_{2q+1}[i]) — the operand pair of an FFMA2 — without register shuffling.  Rows past `cnt` repeat row cnt-1.
template <int D>
__device__ __forceinline__ void sw_load_x_clamped(const float* __restrict__ x, int64_t n0, int cnt, float (&pre)[D], int lane) {
#pragma unroll
    for (int j = 0; j < D; ++j) {
        const int e = lane + 32 * j, p = e / D, i = e - p * D;
        pre[j] = x[(n0 + min(p, cnt - 1)) * D + i];
    }
}
template <int D>
__device__ __forceinline__ void sw_put_x_pairs(float* xs, const float (&pre)[D], int lane) {
#pragma unroll
    for (int j = 0; j < D; ++j) {
        const int e = lane + 32 * j, p = e / D, i = e - p * D;
        xs[(p >> 1) * 2 * D + 2 * i + (p & 1)] = pre[j];
    }
}
template <int D>
__device__ __forceinline__ void sw_get_pair(const float* xs, int q, float (&xa)[D + 1], float (&xb)[D + 1]) {
    const float2* src = reinterpret_cast<const float2*>(xs + q * 2 * D);
#pragma unroll
    for (int i = 0; i < D; ++i) {
        const float2 v = src[i];
        xa[i] = v.x;
        xb[i] = v.y;
    }
    xa[D] = 1.f;
    xb[D] = 1.f;
}

// ---------------------------------------------------------------------------------------------------- S: statistics of a given state
// ASYNC: the r / u rows of 8 points at a time are staged through a per-warp 3-deep cp.async ring (16-byte copies: needs
// K % 4 == 0 and 16-byte aligned r / u), so ~4 KB per warp are in flight while the previous group is accumulated; the plain
// variant loads them straight into registers.
template <int D, bool ASYNC>
__global__ void __launch_bounds__(SW_WARPS * 32, 2)
sweep_stats_kernel(int64_t N, int K, const float* __restrict__ x, const float* __restrict__ r, const float* __restrict__ u,
                   double* __restrict__ stats, const NgTail tail) {
    constexpr int NA = sw_na(D), GP = 8, NST = 3;
    __shared__ double red[NA + 1][32];
    __shared__ __align__(16) float xsm[SW_WARPS][32 * D];
    extern __shared__ __align__(128) float ring_raw[];                // ASYNC: [SW_WARPS][NST][2][GP * 32]
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float* ring_w = ring_raw + (size_t)wib * NST * 2 * GP * 32;
    for (int t = threadIdx.x; t < (NA + 1) * 32; t += blockDim.x) (&red[0][0])[t] = 0.0;
    __syncthreads();
    const int64_t nwarps = (int64_t)gridDim.x * SW_WARPS, gw = (int64_t)blockIdx.x * SW_WARPS + wib;
    const bool kin = lane < K, has_u = u != nullptr;
    const int kl = kin ? lane : 0;
    float* xs = xsm[wib];
    float acc[NA];
#pragma unroll
    for (int e = 0; e < NA; ++e) acc[e] = 0.f;
    float racc = 0.f;
    constexpr int GPR = SW_RUN / GP;                                 // groups per run
    // linear group counter c of this warp -> first point of the group (runs are strided over the warps of the grid)
    auto group_point = [&](int64_t c) -> int64_t { return (gw + (c / GPR) * nwarps) * SW_RUN + (c % GPR) * GP; };
    auto issue = [&](int64_t c) {
        if constexpr (ASYNC) {
            const int64_t n0 = group_point(c);
            float* dst = ring_w + (c % NST) * 2 * GP * 32;
            const int chunks = GP * K / 4;                           // 16-byte chunks of GP consecutive rows
            for (int e = lane; e < chunks; e += 32) {
                const bool ok = n0 + (4 * e) / K < N;
                sw_cp_async16(dst + 4 * e, ok ? r + n0 * K + 4 * e : r, ok);
                if (has_u) sw_cp_async16(dst + GP * 32 + 4 * e, ok ? u + n0 * K + 4 * e : u, ok);
            }
            sw_cp_commit();
        }
    };
    if constexpr (ASYNC) { issue(0); issue(1); }
    int64_t c = 0;
    for (int64_t run = gw; run * SW_RUN < N; run += nwarps) {
        const int64_t r0 = run * SW_RUN, r1 = min(N, r0 + SW_RUN);
        for (int64_t b0 = r0; b0 < r0 + SW_RUN; b0 += 32) {
            const int cnt = (int)max((int64_t)0, min((int64_t)32, r1 - b0));
            if (cnt > 0) {
                float pre[D];
                sw_load_x<D>(x, b0, cnt, pre, lane);
                __syncwarp();
                sw_put_x<D>(xs, pre, lane);
                __syncwarp();
            }
#pragma unroll 1
            for (int p0 = 0; p0 < 32; p0 += GP, ++c) {
                float rv[GP], uv[GP];
                if constexpr (ASYNC) {
                    issue(c + 2);
                    sw_cp_wait<2>();
                    __syncwarp();
                    const float* src = ring_w + (c % NST) * 2 * GP * 32;
#pragma unroll
                    for (int q = 0; q < GP; ++q) {
                        rv[q] = src[q * K + kl];
                        uv[q] = has_u ? src[GP * 32 + q * K + kl] : 1.f;
                    }
                    __syncwarp();
                } else {
#pragma unroll
                    for (int q = 0; q < GP; ++q) {
                        const bool in = p0 + q < cnt;
                        rv[q] = in ? r[(b0 + p0 + q) * K + kl] : 0.f;
                        uv[q] = (in && has_u) ? u[(b0 + p0 + q) * K + kl] : 1.f;
                    }
                }
#pragma unroll
                for (int q = 0; q < GP; ++q) {
                    if (p0 + q < cnt) {
                        float xt[D + 1];
                        sw_get_row<D>(xs, p0 + q, xt);
                        const float rr = kin ? rv[q] : 0.f;
                        racc += rr;
                        sw_accumulate<D>(acc, xt, rr * uv[q]);
                    }
                }
            }
        }
        sw_flush<D>(acc, racc, red, lane);
    }
    if constexpr (ASYNC) sw_cp_wait<0>();
    __syncthreads();
    sw_store_stats<D>(red, K, u != nullptr, stats);
    ng_tail_run<float>(tail, K, D, stats, gridDim.x);
}

// ---------------------------------------------------------------------------------------------------- S (tensor cores): D = 8
// stats[k][f] = sum_n w[n][k] F[n][f] with F = the 45 products of the augmented row [x, 1] is a [K x N] x [N x 48] GEMM: legacy
// warp-level mma.sync.m16n8k8 (tf32 operands split hi/lo, three MMAs per product: hi*hi + lo*hi + hi*lo, ~2^-21 relative per
// term; the running fp32 sums are kept with FADDs and flushed to fp64 every SW_RUN points).  The contraction index is the point: one MMA step covers 8
// points.  A fragment = w^T: the lane (g, t) loads r[n0+t][4g..4g+3] and r[n0+t+4][4g..4g+3] as two 128-bit rows (coalesced:
// the eight g-lanes of a t cover one 128-byte row), output rows are assigned to components so that no shuffle is needed
// (m-tile mt, row g -> component 4g + 2mt, row g+8 -> component 4g + 2mt + 1).  B fragment = the feature (slot 8q + g) of the
// points t, t+4, gathered from the staged x rows with per-lane constant offsets (the constants 1 and 0 sit behind the rows).
// 36 MMAs + ~150 other instructions per 8 points (the FP32 lane <-> component kernel: 600).  What then binds is memory-level
// parallelism: every warp owns a ring of SWM_NST cp.async stages, one stage = the contiguous x, r and u rows of a group of 8
// points, issued three groups ahead; groups are dealt to the warps in contiguous, balanced chunks (runs of 256 points left
// 18 % of the warps idle in the last wave at N = 10^6).
// d = a b (zero accumulator input)
__device__ __forceinline__ void sw_mma_tf32_zero(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}
__device__ __forceinline__ void sw_mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void sw_split_tf32(float v, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(v) & 0xffffe000u;
    lo = __float_as_uint(v - __uint_as_float(hi));        // the tensor core ignores the 13 low mantissa bits of lo
}

// per-warp ring of cp.async stages: one stage = the x, r and u rows of one group of 8 points
constexpr int SWM_NST = 4;                      // stages per warp: 3 groups (6.9 KB) in flight behind the one being multiplied
constexpr int SWM_XA = 128;                     // x area: 64 floats of rows | 1, 0 at [64], [65] and again at [96], [97]
constexpr int SWM_RS = 36;                      // row stride of the staged r / u rows: 4 t-rows x 8 g-chunks hit 32 distinct bank quads
constexpr int SWM_RA = 8 * SWM_RS;              // floats of one r (or u) area
constexpr int SWM_STAGE = SWM_XA + 2 * SWM_RA;  // floats per stage: x area | r rows | u rows
constexpr int SWM_FLUSH = 128;                  // groups between two flushes of the fp32 sums (1024 points; the fp64 shared-memory
                                                // atomics of a flush cost as much as ~15 groups)

__global__ void __launch_bounds__(SW_WARPS * 32, 2)
sweep_stats_mma_kernel(int64_t N, int K, const float* __restrict__ x, const float* __restrict__ r, const float* __restrict__ u,
                       double* __restrict__ stats, const NgTail tail) {
    constexpr int D = 8, NA = sw_na(D), NT = 6;                     // 45 features in 6 n-tiles of 8 slots (3 spare = 0)
    __shared__ double red[NA + 1][32];
    extern __shared__ __align__(128) float ring_raw[];               // [SW_WARPS][SWM_NST][SWM_STAGE]
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    for (int e = threadIdx.x; e < (NA + 1) * 32; e += blockDim.x) (&red[0][0])[e] = 0.0;
    float* ring = ring_raw + (size_t)wib * SWM_NST * SWM_STAGE;
    for (int e = lane; e < SWM_NST * SWM_STAGE; e += 32) {           // the constants behind the x rows (never overwritten)
        const int o = e % SWM_STAGE;
        ring[e] = (o == 64 || o == 96) ? 1.f : 0.f;
    }
    __syncthreads();
    // slot 8q + g -> (i, j) of the augmented lower triangle (e = i (i + 1) / 2 + j); index 8 = the constant 1, spare slots = 0 * 0.
    // Offsets are relative to the row of point t (row t+4 is 32 floats further): the constants sit at [64 + d] and [96 + d].
    int oi[NT], oj[NT];
#pragma unroll
    for (int q = 0; q < NT; ++q) {
        const int e = 8 * q + g;
        int i = 0;
        while ((i + 1) * (i + 2) / 2 <= e) ++i;
        const int j = e - i * (i + 1) / 2;
        oi[q] = e >= NA ? 65 - 8 * t : (i == 8 ? 64 - 8 * t : i);
        oj[q] = e >= NA ? 65 - 8 * t : (j == 8 ? 64 - 8 * t : j);
    }
    // contiguous chunk of groups per warp (balanced to one group)
    const int64_t G = (N + 7) / 8, nwarps = (int64_t)gridDim.x * SW_WARPS, gw = (int64_t)blockIdx.x * SW_WARPS + wib;
    const int64_t per = (G + nwarps - 1) / nwarps, gbeg = min(G, gw * per), gend = min(G, gbeg + per);
    const int64_t ng = gend - gbeg;
    const bool kin = 4 * g < K, has_u = u != nullptr;
    // copies of group gbeg + c into stage c % NST: 16-byte cp.async per lane (rows past N are zero-filled: zero weights, finite
    // x), one commit group per stage.  (Bulk TMA copies of 256 B - 1 KB were request-bound: ~85 cycles each, 119 us per sweep.)
    // this lane's (at most two) 16-byte chunks of the 8 x K block of r (and u): source / destination offsets are per-lane constants
    int csrc[2], cdst[2], crow[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int e = lane + 32 * h, row = (4 * e) / K;
        csrc[h] = 4 * e;
        cdst[h] = SWM_XA + row * SWM_RS + (4 * e - row * K);
        crow[h] = e < 2 * K ? row : 1 << 20;                         // no such chunk: never copied
    }
    auto issue = [&](int64_t c) {
        if (c < ng) {
            const int64_t n0 = (gbeg + c) * 8;
            float* st = ring + (c % SWM_NST) * SWM_STAGE;
            if (lane < 16) sw_cp_async16(st + 4 * lane, x + n0 * D + 4 * lane, n0 + (lane >> 1) < N);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (crow[h] < 8) {
                    const bool ok = n0 + crow[h] < N;
                    sw_cp_async16(st + cdst[h], ok ? r + n0 * K + csrc[h] : r, ok);
                    if (has_u) sw_cp_async16(st + cdst[h] + SWM_RA, ok ? u + n0 * K + csrc[h] : u, ok);
                }
            }
        }
        sw_cp_commit();                                              // (an empty group keeps the wait_group arithmetic uniform)
    };
    for (int64_t c = 0; c < SWM_NST - 1; ++c) issue(c);
    float acc[2][NT][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int q = 0; q < NT; ++q)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[m][q][c] = 0.f;
    float4 racc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f), one4 = make_float4(1.f, 1.f, 1.f, 1.f);
    auto flush = [&]() {
        // fp32 sums -> fp64 shared sums.  c0/c1: component 4g+2m, slots 8q+2t, +1; c2/c3: component 4g+2m+1
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int q = 0; q < NT; ++q) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int slot = 8 * q + 2 * t + (c & 1), comp = 4 * g + 2 * m + (c >> 1);
                    if (slot < NA && comp < K) atomicAdd(&red[slot][comp], (double)acc[m][q][c]);
                    acc[m][q][c] = 0.f;
                }
            }
        // sum r: the four t-lanes of a g hold partial sums of the same four components
        racc.x += __shfl_xor_sync(0xffffffffu, racc.x, 1); racc.x += __shfl_xor_sync(0xffffffffu, racc.x, 2);
        racc.y += __shfl_xor_sync(0xffffffffu, racc.y, 1); racc.y += __shfl_xor_sync(0xffffffffu, racc.y, 2);
        racc.z += __shfl_xor_sync(0xffffffffu, racc.z, 1); racc.z += __shfl_xor_sync(0xffffffffu, racc.z, 2);
        racc.w += __shfl_xor_sync(0xffffffffu, racc.w, 1); racc.w += __shfl_xor_sync(0xffffffffu, racc.w, 2);
        if (t == 0 && kin) {
            atomicAdd(&red[NA][4 * g + 0], (double)racc.x);
            atomicAdd(&red[NA][4 * g + 1], (double)racc.y);
            atomicAdd(&red[NA][4 * g + 2], (double)racc.z);
            atomicAdd(&red[NA][4 * g + 3], (double)racc.w);
        }
        racc = zero4;
    };
#pragma unroll 1
    for (int64_t c = 0; c < ng; ++c) {
        // refill the stage that group c-1 used (every lane is past it: __syncwarp), then wait for group c's copies
        __syncwarp();
        issue(c + SWM_NST - 1);
        sw_cp_wait<SWM_NST - 1>();
        __syncwarp();
        const float* st = ring + (c % SWM_NST) * SWM_STAGE;
        const float4 r_a = kin ? *reinterpret_cast<const float4*>(st + SWM_XA + t * SWM_RS + 4 * g) : zero4;
        const float4 r_b = kin ? *reinterpret_cast<const float4*>(st + SWM_XA + (t + 4) * SWM_RS + 4 * g) : zero4;
        const float4 u_a = (kin && has_u) ? *reinterpret_cast<const float4*>(st + SWM_XA + SWM_RA + t * SWM_RS + 4 * g) : one4;
        const float4 u_b = (kin && has_u) ? *reinterpret_cast<const float4*>(st + SWM_XA + SWM_RA + (t + 4) * SWM_RS + 4 * g) : one4;
        racc.x += r_a.x + r_b.x; racc.y += r_a.y + r_b.y; racc.z += r_a.z + r_b.z; racc.w += r_a.w + r_b.w;
        // A fragments: m-tile 0 = components (4g, 4g+1), m-tile 1 = (4g+2, 4g+3); columns = points t, t+4
        uint32_t ah[2][4], al[2][4];
        sw_split_tf32(r_a.x * u_a.x, ah[0][0], al[0][0]);
        sw_split_tf32(r_a.y * u_a.y, ah[0][1], al[0][1]);
        sw_split_tf32(r_b.x * u_b.x, ah[0][2], al[0][2]);
        sw_split_tf32(r_b.y * u_b.y, ah[0][3], al[0][3]);
        sw_split_tf32(r_a.z * u_a.z, ah[1][0], al[1][0]);
        sw_split_tf32(r_a.w * u_a.w, ah[1][1], al[1][1]);
        sw_split_tf32(r_b.z * u_b.z, ah[1][2], al[1][2]);
        sw_split_tf32(r_b.w * u_b.w, ah[1][3], al[1][3]);
        const float* xa = st + t * D;
        const float* xb = xa + 4 * D;
#pragma unroll
        for (int q = 0; q < NT; ++q) {
            uint32_t bh0, bl0, bh1, bl1;
            sw_split_tf32(xa[oi[q]] * xa[oj[q]], bh0, bl0);
            sw_split_tf32(xb[oi[q]] * xb[oj[q]], bh1, bl1);
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                // the tensor core's fp32 accumulation truncates: only the 8-point partial goes through it, the
                // running sums are added with round-to-nearest FADDs (5e-6 -> 1e-7 relative over a run)
                float d[4];
                sw_mma_tf32_zero(d, al[m], bh0, bh1);
                sw_mma_tf32(d, ah[m], bl0, bl1);
                sw_mma_tf32(d, ah[m], bh0, bh1);
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) acc[m][q][cc] += d[cc];
            }
        }
        if ((c + 1) % SWM_FLUSH == 0 && c + 1 < ng) flush();
    }
    // last flush without atomics: every warp dumps its fp32 sums into the (now idle) ring, 256 threads add the eight partials
    // of each (slot, component) in double — the map (m, q, c, g, t) -> (slot, component) is one-to-one, so the add is plain
    sw_cp_wait<0>();
    __syncthreads();
    float* dump = ring_raw;                                          // [SW_WARPS][2 * NT * 4][32]
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int q = 0; q < NT; ++q)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                dump[((wib * 2 * NT * 4) + (m * NT + q) * 4 + c) * 32 + lane] = acc[m][q][c];
                acc[m][q][c] = 0.f;
            }
    flush();                                                         // (sum r; the accumulators are zero by now)
    __syncthreads();
    for (int idx = threadIdx.x; idx < 2 * NT * 4 * 32; idx += blockDim.x) {
        const int reg = idx >> 5, l = idx & 31, m = reg / (NT * 4), q = (reg >> 2) % NT, c = reg & 3;
        const int slot = 8 * q + 2 * (l & 3) + (c & 1), comp = 4 * (l >> 2) + 2 * m + (c >> 1);
        if (slot < NA && comp < K) {
            double sum = 0.0;
#pragma unroll
            for (int w = 0; w < SW_WARPS; ++w) sum += (double)dump[((w * 2 * NT * 4) + reg) * 32 + l];
            red[slot][comp] += sum;
        }
    }
    __syncthreads();
    sw_store_stats<D>(red, K, u != nullptr, stats);
    ng_tail_run<float>(tail, K, D, stats, gridDim.x);
}

// ---------------------------------------------------------------------------------------------------- E: e-step (+ next statistics)
// Points are processed in groups of GP = 8 (four packed pairs): the eight score chains, the eight soft-max reductions and the
// eight normalisations are independent, so the shuffle / MUFU latencies overlap instead of serialising per point.
template <int D, bool SMM, bool WRITE, bool STATS>
__global__ void __launch_bounds__((STATS ? SW_WARPS_FUSED : SW_WARPS) * 32, STATS ? 1 : 2)
sweep_estep_kernel(int64_t N, int K, const float* __restrict__ x, const float* __restrict__ rec, float* __restrict__ r,
                   float* __restrict__ u, double* __restrict__ stats) {
    constexpr int NP = sw_np(D), RS = sw_rs(D), NA = sw_na(D), WARPS = STATS ? SW_WARPS_FUSED : SW_WARPS, GP = 8;
    __shared__ double red[STATS ? NA + 1 : 1][32];
    __shared__ __align__(16) float xsm[WARPS][32 * D];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if constexpr (STATS) {
        for (int t = threadIdx.x; t < (NA + 1) * 32; t += blockDim.x) (&red[0][0])[t] = 0.0;
        __syncthreads();
    }
    const int64_t nwarps = (int64_t)gridDim.x * WARPS, gw = (int64_t)blockIdx.x * WARPS + wib;
    const bool kin = lane < K;
    float rc[RS];
    {
        const float4* rp = reinterpret_cast<const float4*>(rec + (size_t)(kin ? lane : 0) * RS);
#pragma unroll
        for (int q = 0; q < RS / 4; ++q) {
            const float4 v = rp[q];
            rc[4 * q] = v.x; rc[4 * q + 1] = v.y; rc[4 * q + 2] = v.z; rc[4 * q + 3] = v.w;
        }
    }
    const float cp2 = rc[NP + D], hk2 = rc[NP + D + 1], ek = rc[NP + D + 2], Dk = rc[NP + D + 3];
    float* xs = xsm[wib];
    float acc[STATS ? NA : 1];
#pragma unroll
    for (int e = 0; e < (STATS ? NA : 1); ++e) acc[e] = 0.f;
    float racc = 0.f;
    // every warp owns one contiguous, balanced range of 32-point batches (runs of SW_RUN points dealt round-robin left 18 % of
    // the warps without work in the last wave at N = 10^6); inside it the fp32 statistics are flushed every SW_RUN points
    const int64_t nbatch = (N + 31) / 32, per = (nbatch + nwarps - 1) / nwarps;
    const int64_t wbeg = min(N, gw * per * 32), wend = min(N, wbeg + per * 32);
    for (int64_t r0 = wbeg; r0 < wend; r0 += SW_RUN) {
        const int64_t r1 = min(wend, r0 + SW_RUN);
        float pre[D];
        sw_load_x_clamped<D>(x, r0, (int)min((int64_t)32, r1 - r0), pre, lane);
        for (int64_t b0 = r0; b0 < r1; b0 += 32) {
            // every batch is evaluated as 32 points (rows past the end repeat the last one): no data-dependent branch around
            // the warp collectives; only the stores and the statistics weights are predicated
            const int cnt = (int)min((int64_t)32, r1 - b0);
            __syncwarp();
            sw_put_x_pairs<D>(xs, pre, lane);
            __syncwarp();
            if (b0 + 32 < r1) sw_load_x_clamped<D>(x, b0 + 32, (int)min((int64_t)32, r1 - b0 - 32), pre, lane);   // prefetch
#pragma unroll 1
            for (int p0 = 0; p0 < 32; p0 += GP) {
                // ---- scores of GP points: q = v (x - m)^T P (x - m) in triangular form, two points per packed FFMA2
                float qv[GP];
#pragma unroll
                for (int pp = 0; pp < GP / 2; ++pp) {
                    float xa[D + 1], xb[D + 1], da[D], db[D];
                    sw_get_pair<D>(xs, p0 / 2 + pp, xa, xb);
#pragma unroll
                    for (int i = 0; i < D; ++i) sw_fadd2_bcast(da[i], db[i], xa[i], xb[i], rc[NP + i]);   // x - m (record holds -m)
                    float qa = 0.f, qb = 0.f;
#pragma unroll
                    for (int i = 0; i < D; ++i) {
                        float sa = 0.f, sb = 0.f;
#pragma unroll
                        for (int c2 = 0; c2 < i; ++c2) sw_ffma2_bcast(sa, sb, rc[i * (i + 1) / 2 + c2], da[c2], db[c2]);
                        sw_ffma2_bcast(sa, sb, rc[i * (i + 1) / 2 + i], da[i], db[i]);
                        sw_ffma2(qa, qb, da[i], db[i], sa, sb);
                    }
                    qv[2 * pp] = qa;
                    qv[2 * pp + 1] = qb;
                }
                // ---- soft-max over the components (lanes), base 2: log2 rho = cp2 - hk2 q  (gmm.py:141-151, smm.py:122-128)
                float ev[GP];
#pragma unroll
                for (int g = 0; g < GP; ++g) {
                    const float l2 = kin ? fmaf(-hk2, qv[g], cp2) : -CUDART_INF_F;
                    ev[g] = sw_ex2(l2 - sw_warp_max(l2));
                }
                float sv[GP];
#pragma unroll
                for (int g = 0; g < GP; ++g) sv[g] = ev[g];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                    for (int g = 0; g < GP; ++g) sv[g] += __shfl_xor_sync(0xffffffffu, sv[g], o);
                }
                float rr[GP], uu[GP];
#pragma unroll
                for (int g = 0; g < GP; ++g) {
                    rr[g] = __fdividef(ev[g], sv[g]);
                    uu[g] = SMM ? __fdividef(Dk, qv[g] + ek) : 1.f;                          // smm.py:131-137
                }
                if (WRITE && kin) {
#pragma unroll
                    for (int g = 0; g < GP; ++g) {
                        if (p0 + g < cnt) {
                            r[(b0 + p0 + g) * K + lane] = rr[g];
                            if (SMM) u[(b0 + p0 + g) * K + lane] = uu[g];
                        }
                    }
                }
                if constexpr (STATS) {
#pragma unroll
                    for (int pp = 0; pp < GP / 2; ++pp) {
                        float xa[D + 1], xb[D + 1];
                        sw_get_pair<D>(xs, p0 / 2 + pp, xa, xb);
                        const float ra = p0 + 2 * pp < cnt ? rr[2 * pp] : 0.f, rb = p0 + 2 * pp + 1 < cnt ? rr[2 * pp + 1] : 0.f;
                        racc += ra + rb;
                        sw_accumulate<D>(acc, xa, ra * uu[2 * pp]);
                        sw_accumulate<D>(acc, xb, rb * uu[2 * pp + 1]);
                    }
                }
            }
        }
        if constexpr (STATS) sw_flush<D>(acc, racc, red, lane);
    }
    if constexpr (STATS) {
        __syncthreads();
        sw_store_stats<D>(red, K, SMM, stats);
    }
}

// ---------------------------------------------------------------------------------------------------- host side
static int sw_grid(const void* kern, int64_t N, int warps, size_t dyn_smem = 0) {
    int dev = 0, sms = 148, occ = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, warps * 32, dyn_smem);
    if (occ < 1) occ = 1;
    int64_t grid = (int64_t)sms * occ;
    const int64_t runs = (N + SW_RUN - 1) / SW_RUN, need = (runs + warps - 1) / warps;
    if (grid > need) grid = need;
    return (int)(grid < 1 ? 1 : grid);
}

template <int D>
static int sw_launch_stats(int64_t N, int K, const float* x, const float* r, const float* u, double* stats, const NgTail& tail,
                           cudaStream_t st) {
    const bool aligned = (K % 4 == 0) && ((reinterpret_cast<uintptr_t>(r) | reinterpret_cast<uintptr_t>(u)) % 16 == 0);
    if (D == 8 && aligned && reinterpret_cast<uintptr_t>(x) % 16 == 0) {
        auto kern = sweep_stats_mma_kernel;
        constexpr int ring_bytes = SW_WARPS * SWM_NST * SWM_STAGE * (int)sizeof(float);
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ring_bytes);
        if (e != cudaSuccess) return (int)e;
        int dev = 0, sms = 148, occ = 1;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, SW_WARPS * 32, ring_bytes);
        int64_t grid = (int64_t)sms * (occ < 1 ? 1 : occ);
        const int64_t need = ((N + 7) / 8 + SW_WARPS - 1) / SW_WARPS;     // at least one group of 8 points per warp
        if (grid > need) grid = need;
        kern<<<(unsigned)(grid < 1 ? 1 : grid), SW_WARPS * 32, ring_bytes, st>>>(N, K, x, r, u, stats, tail);
        return launch_status();
    }
    if (aligned) {
        auto kern = sweep_stats_kernel<D, true>;
        constexpr int ring_bytes = SW_WARPS * 3 * 2 * 8 * 32 * (int)sizeof(float);
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ring_bytes);
        if (e != cudaSuccess) return (int)e;
        kern<<<sw_grid((const void*)kern, N, SW_WARPS, ring_bytes), SW_WARPS * 32, ring_bytes, st>>>(N, K, x, r, u, stats, tail);
    } else {
        auto kern = sweep_stats_kernel<D, false>;
        kern<<<sw_grid((const void*)kern, N, SW_WARPS), SW_WARPS * 32, 0, st>>>(N, K, x, r, u, stats, tail);
    }
    return launch_status();
}
template <int D, bool SMM>
static int sw_launch_estep(int64_t N, int K, const float* x, const float* rec, float* r, float* u, double* stats, bool write,
                           cudaStream_t st) {
#define VMP_SW_GO(W, S)                                                                                      \
    {                                                                                                        \
        auto kern = sweep_estep_kernel<D, SMM, W, S>;                                                        \
        constexpr int WP = S ? SW_WARPS_FUSED : SW_WARPS;                                                    \
        kern<<<sw_grid((const void*)kern, N, WP), WP * 32, 0, st>>>(N, K, x, rec, r, u, stats);              \
        return launch_status();                                                                              \
    }
    if (write && stats) VMP_SW_GO(true, true)
    if (write) VMP_SW_GO(true, false)
    VMP_SW_GO(false, true)
#undef VMP_SW_GO
}

// workspace: stats[2][K * stats_len] double | rec[K][RS] float | P_k[K,D,D] T | cst[K] T   (T counted as 8 bytes)
size_t mixture_fit_workspace_bytes(int K, int D) {
    return sizeof(double) * 2 * (size_t)K * stats_len(D) + sizeof(float) * (size_t)K * sw_rs(D <= 8 ? D : 8) +
           sizeof(double) * ((size_t)K * D * D + K + 2) + 64;
}

// general kernels of suffstats.cu / mixtures.cu (fp64, D > 8, K > 32) through their C entry points
static int gen_suffstats(int64_t N, int K, int D, const float* x, const float* r, const float* u, double* stats, void* st) {
    return vmp_suffstats_f32(N, K, D, x, r, 0, u, stats, st);
}
static int gen_suffstats(int64_t N, int K, int D, const double* x, const double* r, const double* u, double* stats, void* st) {
    return vmp_suffstats_f64(N, K, D, x, r, 0, u, stats, st);
}
static int gen_estep(int64_t N, int K, int D, const float* x, const float* a, const float* b, const float* m, const float* P,
                     const float* v, const float* kap, float* r, float* u, float* pi, float* work, void* st) {
    return vmp_mixture_estep_f32(N, K, D, x, a, b, m, P, v, kap, nullptr, r, u, pi, work, st);
}
static int gen_estep(int64_t N, int K, int D, const double* x, const double* a, const double* b, const double* m, const double* P,
                     const double* v, const double* kap, double* r, double* u, double* pi, double* work, void* st) {
    return vmp_mixture_estep_f64(N, K, D, x, a, b, m, P, v, kap, nullptr, r, u, pi, work, st);
}

template <typename T> struct SweepFast {
    static int run(int64_t, int, int, int, const T*, const T*, T*, T*, double*, double*, int, bool, cudaStream_t) { return -100; }
};

template <typename T>
int mixture_fit(int64_t N, int K, int D, int is_smm, int n_sweeps, const T* x, const T* alpha_0, const T* beta_0, const T* m_0,
                const T* C_0, const T* v_0, const T* kappa_k, T* r, T* u, T* alpha_k, T* beta_k, T* m_k, T* C_k, T* v_k, T* x_k,
                T* S_k, T* pi, void* work, size_t work_bytes, void* stream) {
    if (N <= 0 || K <= 0 || n_sweeps < 1) return VMP_E_BADARG;
    if (D < 1 || D > VMP_MAX_D) return VMP_E_BADDIM;
    if (!x || !alpha_0 || !beta_0 || !m_0 || !C_0 || !v_0 || !r || !alpha_k || !beta_k || !m_k || !C_k || !v_k || !x_k || !S_k ||
        !pi || !work)
        return VMP_E_BADARG;
    if (is_smm && (!kappa_k || !u)) return VMP_E_BADARG;
    if (work_bytes < mixture_fit_workspace_bytes(K, D)) return VMP_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t SLK = (size_t)K * stats_len(D);
    double* stats[2] = {static_cast<double*>(work), static_cast<double*>(work) + SLK};
    float* rec = reinterpret_cast<float*>(stats[1] + SLK);             // [K][RS]   (16-byte aligned: 2 K SL doubles precede it)
    T* Pk = reinterpret_cast<T*>(reinterpret_cast<double*>(rec + (size_t)K * sw_rs(D <= 8 ? D : 8)));   // [K,D,D]
    T* cst = Pk + (size_t)K * D * D;                                    // [K]
    const bool fast = std::is_same<T, float>::value && D <= 8 && K <= 32;
    cudaError_t me = cudaMemsetAsync(stats[0], 0, SLK * sizeof(double), st);
    if (me != cudaSuccess) return (int)me;
    const size_t psm = sizeof(double) * (2 * (size_t)D * (D + 1) + 2 * D + 40);
    if (psm > 48 * 1024) {
        me = cudaFuncSetAttribute(sweep_prepare_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm);
        if (me != cudaSuccess) return (int)me;
    }
    // first M-step: statistics of the caller's state
    int rc = fast ? SweepFast<T>::run(N, K, D, 0, x, nullptr, r, is_smm ? u : nullptr, stats[0], nullptr, 0, false, st)
                  : gen_suffstats(N, K, D, x, r, is_smm ? u : nullptr, stats[0], stream);
    if (rc) return rc;
    for (int s = 0; s < n_sweeps; ++s) {
        const bool last = s + 1 == n_sweeps;
        double* cur = stats[s & 1];
        double* nxt = last ? nullptr : stats[(s + 1) & 1];
        sweep_prepare_kernel<T><<<K, 128, psm, st>>>(K, D, is_smm, cur, nxt, alpha_0, beta_0, m_0, C_0, v_0,
                                                     is_smm ? kappa_k : nullptr, alpha_k, beta_k, m_k, C_k, v_k, x_k, S_k, pi,
                                                     fast ? nullptr : Pk, fast ? nullptr : cst, fast ? rec : nullptr);
        if (int e = launch_status()) return e;
        if (fast) {
            rc = SweepFast<T>::run(N, K, D, 1 + (is_smm ? 1 : 0), x, reinterpret_cast<const T*>(rec), r, u, nullptr, nxt, 0, last, st);
            if (rc) return rc;
        } else {
            rc = gen_estep(N, K, D, x, alpha_k, beta_k, m_k, Pk, v_k, is_smm ? kappa_k : nullptr, r, u, pi, cst, stream);
            if (rc) return rc;
            if (!last) {
                rc = gen_suffstats(N, K, D, x, r, is_smm ? u : nullptr, nxt, stream);
                if (rc) return rc;
            }
        }
    }
    return VMP_OK;
}

// mode 0: statistics of (r, u) -> stats_out; mode 1: GMM e-step; mode 2: SMM e-step (rec = packed records; stats_out = next
// sweep's statistics or nullptr; write = store r / u)
template <> struct SweepFast<float> {
    static int run(int64_t N, int K, int D, int mode, const float* x, const float* rec, float* r, float* u, double* stats_in,
                   double* stats_out, int, bool write, cudaStream_t st, const NgTail& tail = ng_tail_none()) {
#define VMP_SW_CASE(DD)                                                                              \
    case DD:                                                                                         \
        if (mode == 0) return sw_launch_stats<DD>(N, K, x, r, u, stats_in, tail, st);                \
        if (mode == 1) return sw_launch_estep<DD, false>(N, K, x, rec, r, u, stats_out, write, st);  \
        return sw_launch_estep<DD, true>(N, K, x, rec, r, u, stats_out, write, st);
        switch (D) {
            VMP_SW_CASE(1) VMP_SW_CASE(2) VMP_SW_CASE(3) VMP_SW_CASE(4) VMP_SW_CASE(5) VMP_SW_CASE(6) VMP_SW_CASE(7) VMP_SW_CASE(8)
            default: return VMP_E_BADDIM;
        }
#undef VMP_SW_CASE
    }
};

// the phases of one sweep as separate calls (multi-rank sweeps all-reduce the statistics between them)
template <typename T>
int mixture_prepare(int K, int D, int is_smm, const double* stats, double* stats_next, const T* alpha_0, const T* beta_0,
                    const T* m_0, const T* C_0, const T* v_0, const T* kappa_k, T* alpha_k, T* beta_k, T* m_k, T* C_k, T* v_k,
                    T* x_k, T* S_k, T* pi, T* P_k, T* cst, float* rec, void* stream) {
    if (K <= 0 || !stats || !alpha_0 || !beta_0 || !m_0 || !C_0 || !v_0 || !alpha_k || !beta_k || !m_k || !C_k || !v_k || !x_k ||
        !S_k || !pi)
        return VMP_E_BADARG;
    if (is_smm && !kappa_k) return VMP_E_BADARG;
    if (D < 1 || D > VMP_MAX_D) return VMP_E_BADDIM;
    const size_t psm = sizeof(double) * (2 * (size_t)D * (D + 1) + 2 * D + 40);
    if (psm > 48 * 1024) {
        cudaError_t me = cudaFuncSetAttribute(sweep_prepare_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm);
        if (me != cudaSuccess) return (int)me;
    }
    sweep_prepare_kernel<T><<<K, 128, psm, (cudaStream_t)stream>>>(K, D, is_smm, stats, stats_next, alpha_0, beta_0, m_0, C_0, v_0,
                                                                   is_smm ? kappa_k : nullptr, alpha_k, beta_k, m_k, C_k, v_k, x_k,
                                                                   S_k, pi, P_k, cst, rec);
    return launch_status();
}

// fp32 D <= 8 K <= 32 statistics of a given state (called from vmp_suffstats_f32 for these shapes)
int sweep_stats_f32(int64_t N, int K, int D, const float* x, const float* r, const float* u, double* stats, const NgTail& tail,
                    cudaStream_t st) {
    if (D > 8 || K > 32) return -100;
    return SweepFast<float>::run(N, K, D, 0, x, nullptr, const_cast<float*>(r), const_cast<float*>(u), stats, nullptr, 0, false, st,
                                 tail);
}

}  // namespace vmp

extern "C" {
int vmp_mixture_prepare_f32(int K, int D, int is_smm, const double* stats, double* stats_next, const float* alpha_0,
                            const float* beta_0, const float* m_0, const float* C_0, const float* v_0, const float* kappa_k,
                            float* alpha_k, float* beta_k, float* m_k, float* C_k, float* v_k, float* x_k, float* S_k, float* pi,
                            float* P_k, float* cst, float* rec, void* stream) {
    return vmp::mixture_prepare<float>(K, D, is_smm, stats, stats_next, alpha_0, beta_0, m_0, C_0, v_0, kappa_k, alpha_k, beta_k, m_k,
                                       C_k, v_k, x_k, S_k, pi, P_k, cst, rec, stream);
}
int vmp_mixture_prepare_f64(int K, int D, int is_smm, const double* stats, double* stats_next, const double* alpha_0,
                            const double* beta_0, const double* m_0, const double* C_0, const double* v_0, const double* kappa_k,
                            double* alpha_k, double* beta_k, double* m_k, double* C_k, double* v_k, double* x_k, double* S_k,
                            double* pi, double* P_k, double* cst, float* rec, void* stream) {
    return vmp::mixture_prepare<double>(K, D, is_smm, stats, stats_next, alpha_0, beta_0, m_0, C_0, v_0, kappa_k, alpha_k, beta_k,
                                        m_k, C_k, v_k, x_k, S_k, pi, P_k, cst, rec, stream);
}
int vmp_mixture_record_len(int D) { return D <= 8 ? vmp::sw_rs(D) : 0; }
int vmp_mixture_estep_fused_f32(int64_t N, int K, int D, int is_smm, const float* x, const float* rec, float* r, float* u,
                                double* stats_next, int write_state, void* stream) {
    if (N < 0 || K <= 0 || K > 32) return VMP_E_BADARG;
    if (D < 1 || D > 8) return VMP_E_BADDIM;
    if (N == 0) return VMP_OK;
    if (!x || !rec || (!write_state && !stats_next) || (write_state && (!r || (is_smm && !u)))) return VMP_E_BADARG;
    return vmp::SweepFast<float>::run(N, K, D, is_smm ? 2 : 1, x, rec, r, u, nullptr, stats_next, 0, write_state != 0,
                                      (cudaStream_t)stream);
}
size_t vmp_mixture_fit_workspace_bytes(int K, int D) { return vmp::mixture_fit_workspace_bytes(K, D); }
int vmp_mixture_fit_f32(int64_t N, int K, int D, int is_smm, int n_sweeps, const float* x, const float* alpha_0,
                        const float* beta_0, const float* m_0, const float* C_0, const float* v_0, const float* kappa_k, float* r,
                        float* u, float* alpha_k, float* beta_k, float* m_k, float* C_k, float* v_k, float* x_k, float* S_k,
                        float* pi, void* workspace, size_t workspace_bytes, void* stream) {
    return vmp::mixture_fit<float>(N, K, D, is_smm, n_sweeps, x, alpha_0, beta_0, m_0, C_0, v_0, kappa_k, r, u, alpha_k, beta_k, m_k,
                                   C_k, v_k, x_k, S_k, pi, workspace, workspace_bytes, stream);
}
int vmp_mixture_fit_f64(int64_t N, int K, int D, int is_smm, int n_sweeps, const double* x, const double* alpha_0,
                        const double* beta_0, const double* m_0, const double* C_0, const double* v_0, const double* kappa_k,
                        double* r, double* u, double* alpha_k, double* beta_k, double* m_k, double* C_k, double* v_k, double* x_k,
                        double* S_k, double* pi, void* workspace, size_t workspace_bytes, void* stream) {
    return vmp::mixture_fit<double>(N, K, D, is_smm, n_sweeps, x, alpha_0, beta_0, m_0, C_0, v_0, kappa_k, r, u, alpha_k, beta_k,
                                    m_k, C_k, v_k, x_k, S_k, pi, workspace, workspace_bytes, stream);
}
}
