// mixtures.cu — E-steps of the standalone Bayesian GMM (gmm.py:154-198) and Student-t mixture (smm.py:140-164).
//
//   m_dist_nk = v_k (x_n - m_k)^T P_k (x_n - m_k) + D / beta_k                      gmm.py:84-94 == smm.py:88-97
//   GMM : log rho = E log pi + 1/2 E log|Lambda| - 1/2 m_dist ;  r = softmax_k       gmm.py:141-151
//         E log|Lambda| = sum_{i<D} psi((v+1+i)/2) + D log 2 + (det P > 1e-20 ? log det P : 0)   gmm.py:117-131
//   SMM : log r ∝ lgamma((D+kap)/2) - lgamma(kap/2) - D/2 log(kap pi) + E log pi + 1/2 E log|Lambda|
//                 - 1/2 (D+kap) m_dist + log kap   (literal precedence of smm.py:122-124) ; u = (D+kap)/(m_dist+kap)
//         E log|Lambda| = sum_{i<D} psi((v+i)/2) + D log 2 + logdet P                 smm.py:100-110
//
// Kernel 1 (K CTAs): per-component constants c_k (double Cholesky of P_k for the log-determinant) and pi_k.
// Kernel 2: one thread per point, components staged through shared memory in tiles, un-normalised scores written
// coalesced through a staging tile, online log-sum-exp per point, then a coalesced normalisation pass over the
// CTA's contiguous [128 x K] block (re-read from L2).
#include "common.cuh"

namespace vmp {

constexpr int ES_THREADS = 128;

template <typename T>
__global__ void __launch_bounds__(128)
estep_consts_kernel(int K, int D, const T* __restrict__ alpha_k, const T* __restrict__ P_k, const T* __restrict__ v_k,
                    const T* __restrict__ kappa_k, T* __restrict__ cst, T* __restrict__ pi) {
    extern __shared__ double sm[];
    const int ld = D + 1;
    double* C = sm;
    double* red = C + D * ld;
    const int k = blockIdx.x;
    for (int e = threadIdx.x; e < D * D; e += blockDim.x) C[(e / D) * ld + e % D] = (double)P_k[(size_t)k * D * D + e];
    __syncthreads();
    // in-place lower Cholesky (same routine as prepare.cu, duplicated here to keep the TU self-contained)
    for (int j = 0; j < D; ++j) {
        __syncthreads();
        const double djj = sqrt(C[j * ld + j]);
        __syncthreads();
        if (threadIdx.x == 0) C[j * ld + j] = djj;
        const double inv = 1.0 / djj;
        for (int i = j + 1 + threadIdx.x; i < D; i += blockDim.x) C[i * ld + j] *= inv;
        __syncthreads();
        const int m = D - j - 1;
        for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
            const int i = j + 1 + e / m, c = j + 1 + e % m;
            if (c <= i) C[i * ld + c] -= C[i * ld + j] * C[c * ld + j];
        }
    }
    __syncthreads();
    double sa = 0.0;
    for (int j = threadIdx.x; j < K; j += blockDim.x) sa += (double)alpha_k[j];
    sa = block_sum(sa, red);
    if (threadIdx.x == 0) {
        double logdet = 0.0;
        for (int i = 0; i < D; ++i) logdet += 2.0 * log(C[i * ld + i]);
        const double v = (double)v_k[k];
        const double elogpi = digamma_pos((double)alpha_k[k]) - digamma_pos(sa);
        double sd = 0.0, c;
        if (kappa_k == nullptr) {
            for (int i = 0; i < D; ++i) sd += digamma_pos(0.5 * (v + 1.0 + i));
            // det > 1e-20 guard (gmm.py:120-121); a non-PD P gives NaN here exactly as log(det) would misbehave
            const double ld_guard = (logdet > -46.051701859880914) ? logdet : 0.0;
            c = elogpi + 0.5 * (sd + D * VMP_LOG_2 + ld_guard);
        } else {
            for (int i = 0; i < D; ++i) sd += digamma_pos(0.5 * (v + i));
            const double kap = (double)kappa_k[k];
            c = lgamma(0.5 * (D + kap)) - lgamma(0.5 * kap) - 0.5 * D * log(kap * 3.14159265358979323846) + elogpi +
                0.5 * (sd + D * VMP_LOG_2 + logdet) + log(kap);
        }
        cst[k] = (T)c;
        pi[k] = (T)exp(elogpi);
    }
}

// DT > 0: compile-time D (registers, fully unrolled quadratic form, symmetric: D(D+1)/2 FMAs); DT == 0: run-time D.
// When all K components fit one shared-memory tile (the usual case) the row is normalised in the staging tile and
// r / u are written exactly once, coalesced; otherwise un-normalised scores are written per tile and a second
// coalesced pass over the CTA's contiguous [128 x K] block (an L2 hit) normalises them.
template <typename T, int DT>
__global__ void __launch_bounds__(ES_THREADS)
estep_kernel(int64_t N, int K, int Drt, int KT, const T* __restrict__ x, const T* __restrict__ beta_k,
             const T* __restrict__ m_k, const T* __restrict__ P_k, const T* __restrict__ v_k,
             const T* __restrict__ kappa_k, const uint8_t* __restrict__ mask, const T* __restrict__ cst,
             T* __restrict__ r, T* __restrict__ u_out) {
    constexpr int DM = DT ? DT : VMP_MAX_D;
    const int D = DT ? DT : Drt;
    // DT > 0: one packed record per component, RS scalars, read back as 128-bit broadcasts:
    //   tri[NP] (lower triangle, strictly-lower entries hold P_ic + P_ci) | m[D] | v, D/beta, cst, kappa
    constexpr int NP = DM * (DM + 1) / 2;
    constexpr int RS = DT ? ((NP + DT + 4 + 3) / 4) * 4 : 0;
    extern __shared__ __align__(16) unsigned char smraw[];
    T* Ps = reinterpret_cast<T*>(smraw);                 // DT ? [KT][RS] : [KT][D*D]
    T* ms = Ps + (DT ? (size_t)KT * RS : (size_t)KT * D * D);   // [KT][D]      (run-time D only)
    T* ks = ms + (DT ? 0 : (size_t)KT * D);              // [KT][4]  v_k, D/beta_k, cst, kappa (run-time D only)
    T* stage = ks + (DT ? 0 : (size_t)KT * 4);           // [ES_THREADS][KT+1]
    T* stage2 = stage + (size_t)ES_THREADS * (KT + 1);   // [ES_THREADS][KT+1] (SMM: m_dist -> u)
    __shared__ T lse_s[ES_THREADS];
    const int tid = threadIdx.x;
    const int64_t n0 = (int64_t)blockIdx.x * ES_THREADS;
    const int64_t n = n0 + tid;
    const int np = (int)min((int64_t)ES_THREADS, N - n0);
    const bool smm = kappa_k != nullptr;
    const bool single = KT >= K;

    T xv[DM];
    bool mk[DM];
#pragma unroll
    for (int i = 0; i < D; ++i) {
        xv[i] = n < N ? x[n * D + i] : T(0);
        mk[i] = (mask != nullptr && n < N) ? (mask[n * D + i] != 0) : false;
    }
    T run_max = -CUDART_INF_F;
    double run_sum = 0.0;

    for (int kt0 = 0; kt0 < K; kt0 += KT) {
        const int kn = min(KT, K - kt0);
        __syncthreads();
        if (DT) {
            for (int e = tid; e < kn * RS; e += blockDim.x) {
                const int kk = e / RS, o = e - kk * RS, k = kt0 + kk;
                const T* Pk = P_k + (size_t)k * D * D;
                T v = T(0);
                if (o < NP) {
                    int i = 0;
                    while ((i + 1) * (i + 2) / 2 <= o) ++i;
                    const int c = o - i * (i + 1) / 2;
                    v = c < i ? Pk[i * D + c] + Pk[c * D + i] : Pk[i * D + i];
                } else if (o < NP + D) {
                    v = m_k[(size_t)k * D + (o - NP)];
                } else if (o == NP + D) {
                    v = v_k[k];
                } else if (o == NP + D + 1) {
                    v = T(D) / beta_k[k];
                } else if (o == NP + D + 2) {
                    v = cst[k];
                } else if (o == NP + D + 3) {
                    v = smm ? kappa_k[k] : T(0);
                }
                Ps[e] = v;
            }
        } else {
            for (int e = tid; e < kn * D * D; e += blockDim.x) Ps[e] = P_k[(size_t)kt0 * D * D + e];
            for (int e = tid; e < kn * D; e += blockDim.x) ms[e] = m_k[(size_t)kt0 * D + e];
            for (int e = tid; e < kn; e += blockDim.x) {
                ks[4 * e + 0] = v_k[kt0 + e];
                ks[4 * e + 1] = T(D) / beta_k[kt0 + e];
                ks[4 * e + 2] = cst[kt0 + e];
                ks[4 * e + 3] = smm ? kappa_k[kt0 + e] : T(0);
            }
        }
        __syncthreads();
        if (n < N) {
            for (int kk = 0; kk < kn; ++kk) {
                T dv[DM];
                T q = T(0), kv, kdb, kc, kkap;
                if (DT) {
                    T rec[RS ? RS : 1];
                    const T* rp = Ps + (size_t)kk * RS;
#pragma unroll
                    for (int o = 0; o < RS; ++o) rec[o] = rp[o];           // constant offsets: 128-bit broadcasts
#pragma unroll
                    for (int i = 0; i < D; ++i) dv[i] = mk[i] ? T(0) : xv[i] - rec[NP + i];      // gmm.py:106-108
                    // triangular form: q = sum_i d_i (P_ii d_i + sum_{c<i} (P_ic + P_ci) d_c)
#pragma unroll
                    for (int i = 0; i < D; ++i) {
                        T s = T(0);
#pragma unroll
                        for (int c = 0; c < i; ++c) s = fma(rec[i * (i + 1) / 2 + c], dv[c], s);
                        q = fma(dv[i], fma(rec[i * (i + 1) / 2 + i], dv[i], s), q);
                    }
                    kv = rec[NP + D]; kdb = rec[NP + D + 1]; kc = rec[NP + D + 2]; kkap = rec[NP + D + 3];
                } else {
                    const T* P = Ps + (size_t)kk * D * D;
                    const T* m = ms + (size_t)kk * D;
                    for (int i = 0; i < D; ++i) dv[i] = mk[i] ? T(0) : xv[i] - m[i];
                    for (int i = 0; i < D; ++i) {
                        T s = T(0);
                        for (int c = 0; c < D; ++c) s = fma(P[i * D + c], dv[c], s);
                        q = fma(dv[i], s, q);
                    }
                    kv = ks[4 * kk]; kdb = ks[4 * kk + 1]; kc = ks[4 * kk + 2]; kkap = ks[4 * kk + 3];
                }
                const T md = fma(kv, q, kdb);
                const T lr = smm ? kc - T(0.5) * (T(D) + kkap) * md : kc - T(0.5) * md;
                if (smm) stage2[tid * (KT + 1) + kk] = single ? (T(D) + kkap) / (md + kkap) : md;
                stage[tid * (KT + 1) + kk] = lr;
                if (lr > run_max) {
                    run_sum = run_sum * (double)t_exp(run_max - lr) + 1.0;     // T-precision exp, double accumulation
                    run_max = lr;
                } else {
                    run_sum += (double)t_exp(lr - run_max);
                }
            }
            if (single) {      // normalise this thread's row in the staging tile: r, u leave the SM exactly once
                const T lse = run_max + (T)log(run_sum);
                for (int kk = 0; kk < kn; ++kk) stage[tid * (KT + 1) + kk] = t_exp(stage[tid * (KT + 1) + kk] - lse);
            }
        }
        __syncthreads();
        // coalesced write of the tile: rows of the CTA block are contiguous in r
        for (int e = tid; e < np * kn; e += blockDim.x) {
            const int p = e / kn, kk = e - p * kn;
            r[(n0 + p) * K + kt0 + kk] = stage[p * (KT + 1) + kk];
            if (smm) u_out[(n0 + p) * K + kt0 + kk] = stage2[p * (KT + 1) + kk];
        }
    }
    if (single) return;
    lse_s[tid] = (n < N) ? run_max + (T)log(run_sum) : T(0);
    __syncthreads();
    // normalisation pass over the contiguous [np x K] block
    for (int e = tid; e < np * K; e += blockDim.x) {
        const int p = e / K, k = e - p * K;
        const size_t o = (size_t)n0 * K + e;
        r[o] = t_exp(r[o] - lse_s[p]);
        if (smm) u_out[o] = (T(D) + kappa_k[k]) / (u_out[o] + kappa_k[k]);           // smm.py:131-137
    }
}

template <typename T, int DT>
static int launch_estep(int64_t N, int K, int D, int KT, size_t smem, const T* x, const T* beta_k, const T* m_k,
                        const T* P_k, const T* v_k, const T* kappa_k, const uint8_t* mask, const T* work, T* r, T* u_out,
                        cudaStream_t st) {
    auto kern = estep_kernel<T, DT>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    const int64_t grid = (N + ES_THREADS - 1) / ES_THREADS;
    if (grid > 0x7fffffffLL) return VMP_E_BADARG;
    kern<<<(unsigned)grid, ES_THREADS, smem, st>>>(N, K, D, KT, x, beta_k, m_k, P_k, v_k, kappa_k, mask, work, r, u_out);
    return launch_status();
}

template <typename T>
int mixture_estep(int64_t N, int K, int D, const T* x, const T* alpha_k, const T* beta_k, const T* m_k, const T* P_k,
                  const T* v_k, const T* kappa_k, const uint8_t* mask, T* r, T* u_out, T* pi, T* work, void* stream) {
    if (N < 0 || K <= 0 || !alpha_k || !beta_k || !m_k || !P_k || !v_k || !pi || !work) return VMP_E_BADARG;
    if (N > 0 && (!x || !r)) return VMP_E_BADARG;
    if (kappa_k != nullptr && ((N > 0 && u_out == nullptr) || mask != nullptr)) return VMP_E_BADARG;
    if (D < 1 || D > VMP_MAX_D) return VMP_E_BADDIM;
    cudaStream_t st = (cudaStream_t)stream;
    {
        const size_t sm = sizeof(double) * ((size_t)D * (D + 1) + 64);
        auto kc = estep_consts_kernel<T>;
        if (sm > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
            if (e != cudaSuccess) return (int)e;
        }
        kc<<<K, 128, sm, st>>>(K, D, alpha_k, P_k, v_k, kappa_k, work, pi);
        if (int e = launch_status()) return e;
    }
    if (N == 0) return VMP_OK;
    // components per shared-memory tile: as many as fit in ~96 KB together with the staging tiles
    const size_t rs = D <= 8 ? (size_t)((D * (D + 1) / 2 + D + 4 + 3) / 4) * 4 : (size_t)D * D + D + 4;
    const size_t per_k = sizeof(T) * (rs + 2 * ES_THREADS);
    int KT = (int)((96 * 1024 - 2 * ES_THREADS * sizeof(T)) / per_k);
    if (KT < 1) KT = 1;
    if (KT > K) KT = K;
    const size_t smem = sizeof(T) * ((size_t)KT * rs + 2 * (size_t)ES_THREADS * (KT + 1)) + 16;
#define VMP_ES(DD) \
    case DD: return launch_estep<T, DD>(N, K, D, KT, smem, x, beta_k, m_k, P_k, v_k, kappa_k, mask, work, r, u_out, st)
    switch (D) {
        VMP_ES(1); VMP_ES(2); VMP_ES(3); VMP_ES(4); VMP_ES(5); VMP_ES(6); VMP_ES(7); VMP_ES(8);
        default: return launch_estep<T, 0>(N, K, D, KT, smem, x, beta_k, m_k, P_k, v_k, kappa_k, mask, work, r, u_out, st);
    }
#undef VMP_ES
}

}  // namespace vmp

extern "C" {
int vmp_mixture_estep_f32(int64_t N, int K, int D, const float* x, const float* alpha_k, const float* beta_k,
                          const float* m_k, const float* P_k, const float* v_k, const float* kappa_k,
                          const uint8_t* missing_mask, float* r, float* u_out, float* pi, float* work, void* stream) {
    return vmp::mixture_estep<float>(N, K, D, x, alpha_k, beta_k, m_k, P_k, v_k, kappa_k, missing_mask, r, u_out, pi,
                                     work, stream);
}
int vmp_mixture_estep_f64(int64_t N, int K, int D, const double* x, const double* alpha_k, const double* beta_k,
                          const double* m_k, const double* P_k, const double* v_k, const double* kappa_k,
                          const uint8_t* missing_mask, double* r, double* u_out, double* pi, double* work,
                          void* stream) {
    return vmp::mixture_estep<double>(N, K, D, x, alpha_k, beta_k, m_k, P_k, v_k, kappa_k, missing_mask, r, u_out, pi,
                                      work, stream);
}
}
