// small_step.cu — the WHOLE local-VMP + natural-gradient step in ONE launch for the launch-bound configurations
// (BASELINE C1: N=100, K=10, D=2, S=10; C2: N=274, K=10, D=6, S=10 — a few thousand (point, component) pairs).
//
// The multi-launch path spends 6-7 launches (phi / theta prologues, local step, selection, statistics, update) of a few
// microseconds of work each.  Here one thread-block CLUSTER (1-8 CTAs on neighbouring SMs, hardware cluster barrier, distributed
// shared memory) runs all phases back to back:
//   0  every CTA builds the K phi records (svae.unpack_recognition_gmm, svae.py:342-358) and the K theta records
//      (niw.natural_to_standard / expected_values / dirichlet.expected_log_pi as used by compute_elbo, svae.py:204-208; or
//      svae.unpack_smm + the Student-t constants, svae.py:361-373, student_t.py:7-39) in its own shared memory, one thread per
//      component, double arithmetic — redundant across the CTAs (K is ~10), which saves a cluster barrier;
//   1  thread per (point, component) pair: Cholesky of P2_k + diag(p1_n), both forward solves, samples, ELBO terms
//      (pair_math.cuh: the same arithmetic as local_step.cu); per-point log-sum-exp -> log r;
//   2  thread per point: Gumbel-max draw of z_n and the selected sample x[n, z_n, 0] (kept from phase 1 in shared memory);
//   3  the CTA's partial statistics [N_k, sum r x, sum r x x^T] in shared memory (thread per statistic entry);
//      cluster barrier;
//   4  every CTA reduces a slice of the statistics over the cluster through DSMEM, stores them, and applies
//      theta <- (1 - rho) theta + rho (prior + statistics)  (svae.m_step 154-176 + update_gmm_params 376-403) to that slice;
//      CTA 0 adds up the ELBO partials.
// Single GPU only (several ranks need the all-reduce between 3 and 4).  D <= 8, K <= 32, N*K <= 8192, fp32 and fp64.
#include <cooperative_groups.h>

#include "common.cuh"
#include "pair_math.cuh"

namespace cg = cooperative_groups;

namespace vmp {

constexpr int SM_THREADS_F32 = 512;         // fp32: <= 128 registers per thread
constexpr int SM_THREADS_F64 = 256;
template <typename T> struct SmThreads { static constexpr int value = SM_THREADS_F64; };
template <> struct SmThreads<float> { static constexpr int value = SM_THREADS_F32; };
constexpr int SM_MAX_PAIRS_PER_CTA = 1024;
constexpr int SM_MAX_K = 32;
#ifndef SM_MAX_CLUSTER
#define SM_MAX_CLUSTER 16
#endif

template <typename T> struct SmallStepParams {
    int N, K, S, den_mode, only_alpha, ppc, split;   // ppc: points per CTA; split: threads sharing the S samples of a pair
    const T *eta1, *eta2d;
    const T *eta1_phi2, *L_raw, *pi_raw;          // phi_gmm
    const T *th0, *th1, *th2, *th3, *th4;         // theta: (alpha, A, b, beta, v_hat) or (alpha, mu, L_raw, dof, -)
    const T *p0, *p1, *p2, *p3, *p4;              // prior (alpha, A, b, beta, v_hat)
    T *u0, *u1, *u2, *u3, *u4;                    // theta to update in place (== th* for the Gaussian model)
    const T *noise, *gum_u;
    uint64_t seed, pair_offset;
    double rho;
    const double* rho_dev;
    T *log_r, *x_sample, *x_k_samples;
    int32_t* z;
    double *stats, *elbo_acc;
};

// ---- K-sized prologues, one thread per component, double arithmetic in local arrays (D <= 8) -------------------------------
template <typename T, int D>
__device__ void small_phi_record(int k, int K, const T* __restrict__ eta1_phi2, const T* __restrict__ L_raw,
                                 const T* __restrict__ pi_raw, T* __restrict__ out) {
    double L[D * D], vec[D];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
            const double v = (double)L_raw[(size_t)k * D * D + i * D + j];
            L[i * D + j] = j < i ? v : (j == i ? t_softplus<double>(v) : 0.0);          // svae.py:349-350
        }
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
            double s = 0.0;
#pragma unroll
            for (int c = 0; c < D; ++c) s += (c <= i && c <= j) ? L[i * D + c] * L[j * D + c] : 0.0;
            out[i * D + j] = (T)s;                                                        // P2 = L L^T
        }
    double ld2 = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) {
        double s = (double)eta1_phi2[(size_t)k * D + i];
#pragma unroll
        for (int c = 0; c < i; ++c) s -= L[i * D + c] * vec[c];
        vec[i] = s / L[i * D + i];
        ld2 += log(L[i * D + i]);
    }
#pragma unroll
    for (int ii = 0; ii < D; ++ii) {
        const int i = D - 1 - ii;
        double s = vec[i];
#pragma unroll
        for (int c = i + 1; c < D; ++c) s -= L[c * D + i] * vec[c];
        vec[i] = s / L[i * D + i];
    }
#pragma unroll
    for (int i = 0; i < D; ++i) {
        out[D * D + i] = (T)vec[i];                                                       // mu2 = P2^-1 eta1
        out[D * D + D + i] = eta1_phi2[(size_t)k * D + i];
    }
    double mx = -CUDART_INF;
    for (int j = 0; j < K; ++j) mx = fmax(mx, (double)pi_raw[j]);
    double se = 0.0;
    for (int j = 0; j < K; ++j) se += exp((double)pi_raw[j] - mx);
    out[D * D + 2 * D] = (T)((double)pi_raw[k] - mx - log(se));                           // log softmax (svae.py:356)
    out[D * D + 2 * D + 1] = (T)(2.0 * ld2);
    out[D * D + 2 * D + 2] = T(0);
    out[D * D + 2 * D + 3] = T(0);
}

template <int D> __device__ __forceinline__ void small_chol(double* C) {                  // in place, lower
#pragma unroll
    for (int j = 0; j < D; ++j) {
        double s = C[j * D + j];
#pragma unroll
        for (int c = 0; c < j; ++c) s -= C[j * D + c] * C[j * D + c];
        const double d = sqrt(s);
        C[j * D + j] = d;
#pragma unroll
        for (int i = j + 1; i < D; ++i) {
            double t = C[i * D + j];
#pragma unroll
            for (int c = 0; c < j; ++c) t -= C[i * D + c] * C[j * D + c];
            C[i * D + j] = t / d;
        }
    }
}
template <int D> __device__ __forceinline__ void small_tri_inverse(const double* L, double* W) {
#pragma unroll
    for (int j = 0; j < D; ++j) {
#pragma unroll
        for (int i = 0; i < D; ++i) {
            if (i < j) { W[i * D + j] = 0.0; continue; }
            if (i == j) { W[i * D + j] = 1.0 / L[j * D + j]; continue; }
            double s = 0.0;
#pragma unroll
            for (int c = 0; c < D; ++c) s += (c >= j && c < i) ? L[i * D + c] * W[c * D + j] : 0.0;
            W[i * D + j] = -s / L[i * D + i];
        }
    }
}

template <typename T, int D>
__device__ void small_theta_record_gauss(int k, int K, const T* __restrict__ alpha, const T* __restrict__ A,
                                         const T* __restrict__ b, const T* __restrict__ beta, const T* __restrict__ v_hat,
                                         T* __restrict__ out) {
    double C[D * D], W[D * D];
    const double bk = (double)beta[k], v = (double)v_hat[k] - D - 2.0;                    // niw.py:42
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j)
            C[i * D + j] = (double)A[(size_t)k * D * D + i * D + j] -
                           (double)b[(size_t)k * D + i] * ((double)b[(size_t)k * D + j] / bk);   // niw.py:35-41
    small_chol<D>(C);
    small_tri_inverse<D>(C, W);
    const double sv = sqrt(v);                                                            // E[Sigma] = C / v (niw.py:8-17)
    double hl = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) {
#pragma unroll
        for (int j = 0; j < D; ++j) out[i * D + j] = (T)(j <= i ? sv * W[i * D + j] : 0.0);
        out[D * D + i] = (T)((double)b[(size_t)k * D + i] / bk);
        hl += log(C[i * D + i]);
    }
    double sa = 0.0;
    for (int j = 0; j < K; ++j) sa += (double)alpha[j] + 1.0;
    const double logdetP = D * log(v) - 2.0 * hl;
    const double elogpi = digamma_pos((double)alpha[k] + 1.0) - digamma_pos(sa);          // dirichlet.py:8-12
    out[D * D + D + 0] = (T)(0.5 * logdetP - 0.5 * D * VMP_LOG_2PI + elogpi);
    out[D * D + D + 1] = T(0);
    out[D * D + D + 2] = (T)elogpi;
    out[D * D + D + 3] = (T)logdetP;
}

template <typename T, int D>
__device__ void small_theta_record_student(int k, int K, const T* __restrict__ alpha, const T* __restrict__ mu,
                                           const T* __restrict__ L_raw, const T* __restrict__ dof, T* __restrict__ out) {
    double L[D * D], W[D * D];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
            const double v = (double)L_raw[(size_t)k * D * D + i * D + j];
            L[i * D + j] = j < i ? v : (j == i ? t_softplus<double>(v) : 0.0);          // svae.py:365-371
        }
    small_tri_inverse<D>(L, W);
    double hl = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) {
#pragma unroll
        for (int j = 0; j < D; ++j) out[i * D + j] = (T)(j <= i ? W[i * D + j] : 0.0);
        out[D * D + i] = mu[(size_t)k * D + i];
        hl += log(L[i * D + i]);
    }
    double sa = 0.0;
    for (int j = 0; j < K; ++j) sa += (double)alpha[j] + 1.0;
    const double nu = (double)dof[k];
    const double elogpi = digamma_pos((double)alpha[k] + 1.0) - digamma_pos(sa);
    out[D * D + D + 0] = (T)(lgamma(0.5 * (nu + D)) - lgamma(0.5 * nu) - 0.5 * D * (VMP_LOG_PI + log(nu)) - hl + elogpi);
    out[D * D + D + 1] = (T)nu;
    out[D * D + D + 2] = (T)elogpi;
    out[D * D + D + 3] = (T)(-2.0 * hl);
}

// ---- the step -------------------------------------------------------------------------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(SmThreads<T>::value) svae_small_step_kernel(const SmallStepParams<T> p) {
    using PM = PairMath<T, D>;
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks(), crank = (int)cluster.block_rank();
    constexpr int PL = D * D + 2 * D + 4, TL = D * D + D + 4, SL = D * D + D + 2;
    const int K = p.K, S = p.S, ppc = p.ppc;
    extern __shared__ __align__(16) unsigned char smraw[];
    double* sstat = reinterpret_cast<double*>(smraw);                 // [K][SL]   this CTA's partial statistics
    double* selbo = sstat + (size_t)K * SL;                           // [4]       this CTA's ELBO partials
    T* prec = reinterpret_cast<T*>(selbo + 4);                        // [K][PL]
    T* trec = prec + (size_t)K * PL;                                  // [K][TL]
    T* sc = trec + (size_t)K * TL;                                    // [ppc*K]   score -> log r
    T* tnum = sc + (size_t)ppc * K;                                   // [ppc*K]
    T* tden = tnum + (size_t)ppc * K;                                 // [ppc*K]
    T* x0 = tden + (size_t)ppc * K;                                   // [ppc*K][D] first sample of every pair
    T* xs = x0 + (size_t)ppc * K * D;                                 // [ppc][D]   selected samples
    T* gum = xs + (size_t)ppc * D;                                    // [ppc*K]   Gumbel noise of the categorical draw
    T* rr = gum + (size_t)ppc * K;                                    // [ppc*K]   responsibilities
    __shared__ double red[32];
    const int tid = threadIdx.x;

    // ---------------- phase 0: per-component records (redundant in every CTA)
    if (tid < K) {
        small_phi_record<T, D>(tid, K, p.eta1_phi2, p.L_raw, p.pi_raw, prec + (size_t)tid * PL);
    } else if (tid >= 32 && tid < 32 + K) {                           // a different warp: both prologues run concurrently
        const int k = tid - 32;
        if (p.den_mode == VMP_DEN_GAUSS) small_theta_record_gauss<T, D>(k, K, p.th0, p.th1, p.th2, p.th3, p.th4, trec + (size_t)k * TL);
        else small_theta_record_student<T, D>(k, K, p.th0, p.th1, p.th2, p.th3, trec + (size_t)k * TL);
    }
    for (int e = tid; e < K * SL + 4; e += blockDim.x) sstat[e] = 0.0;
    const int pt0 = crank * ppc;
    const int npts = max(0, min(ppc, p.N - pt0));
    const int npairs = npts * K;
    for (int e = tid; e < npairs; e += blockDim.x) { tnum[e] = T(0); tden[e] = T(0); }
    __syncthreads();

    // ---------------- phase 1: pairs; `split` threads share the S samples of a pair (each refactors the tiny system: cheaper
    // than waiting for one thread to walk through all samples)
    const int split = p.split;
    int nbad = 0;
    for (int it = tid; it < npairs * split; it += blockDim.x) {
        const int q = it / split, sp = it - q * split;
        const int pl = q / K, k = q - pl * K;
        const int64_t n = pt0 + pl;
        PM pm;
        pm.factor(D, p.eta1 + n * D, p.eta2d + n * D, prec + (size_t)k * PL);
        if (sp == 0) nbad += pm.bad;
        const T* tr = trec + (size_t)k * TL;
        const T cden = tr[D * D + D], nu = tr[D * D + D + 1];
        const uint64_t pair = (uint64_t)n * K + k, gpair = pair + p.pair_offset;
        T snum = T(0), sden = T(0);
        for (int s = sp; s < S; s += split) {
            T eps[D], x[D];
            if (p.noise != nullptr) {
#pragma unroll
                for (int i = 0; i < D; ++i) eps[i] = p.noise[(pair * D + i) * (uint64_t)S + s];
            } else {
#pragma unroll
                for (int qd = 0; qd < (D + 3) / 4; ++qd) {
                    const float4 v = philox_normal4(p.seed, gpair, (uint32_t)s, (uint32_t)qd);
                    if (4 * qd < D) eps[4 * qd] = (T)v.x;
                    if (4 * qd + 1 < D) eps[4 * qd + 1] = (T)v.y;
                    if (4 * qd + 2 < D) eps[4 * qd + 2] = (T)v.z;
                    if (4 * qd + 3 < D) eps[4 * qd + 3] = (T)v.w;
                }
            }
            pm.sample(D, eps, x);
            T e2 = T(0);
#pragma unroll
            for (int i = 0; i < D; ++i) e2 = fma(eps[i], eps[i], e2);
            if (p.x_k_samples != nullptr) {
#pragma unroll
                for (int i = 0; i < D; ++i) p.x_k_samples[((pair * S) + s) * (uint64_t)D + i] = x[i];
            }
            if (s == 0) {
#pragma unroll
                for (int i = 0; i < D; ++i) x0[(size_t)q * D + i] = x[i];
            }
            snum += T(-0.5) * e2;
            sden += den_logprob<T>(p.den_mode, D, PM::maha(D, tr, x), cden, nu);
        }
        if (sp == 0) {
            sc[q] = pm.score;
            snum += T(S) * (pm.hld - T(0.5 * VMP_LOG_2PI) * T(D));
            const T u = p.gum_u != nullptr ? p.gum_u[pair] : (T)philox_uniform_pair(p.seed, gpair);
            gum[q] = gumbel_from_uniform<T>(u);                      // the draw's noise is per pair: no need to serialise it per point
        }
        atomicAdd(&tnum[q], snum / T(S));
        atomicAdd(&tden[q], sden / T(S));
    }
    __syncthreads();
    // ---------------- per-point log-sum-exp, Gumbel-max selection (phase 2), log r / z / x_sample out
    for (int pl = tid; pl < npts; pl += blockDim.x) {
        T* s = sc + (size_t)pl * K;
        T mx = s[0];
        for (int k = 1; k < K; ++k) mx = max(mx, s[k]);
        double se = 0.0;
        for (int k = 0; k < K; ++k) se += (double)t_exp(s[k] - mx);
        const T lse = mx + (T)log(se);
        const int64_t n = pt0 + pl;
        int zb = 0;
        T best = -CUDART_INF_F;
        for (int k = 0; k < K; ++k) {
            const T lr = s[k] - lse;
            s[k] = lr;
            p.log_r[n * K + k] = lr;
            rr[(size_t)pl * K + k] = t_exp(lr);
            const T cand = lr + gum[(size_t)pl * K + k];
            if (cand > best) { best = cand; zb = k; }
        }
        if (p.z != nullptr) p.z[n] = zb;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const T v = x0[((size_t)pl * K + zb) * D + i];
            xs[(size_t)pl * D + i] = v;
            if (p.x_sample != nullptr) p.x_sample[n * D + i] = v;
        }
    }
    __syncthreads();
    // ELBO partials of this CTA
    double e_num = 0.0, e_den = 0.0;
    for (int q = tid; q < npairs; q += blockDim.x) {
        const double r = (double)rr[q];
        e_num += r * ((double)tnum[q] + (double)sc[q]);
        e_den += r * (double)tden[q];
    }
    const double bn = block_sum(e_num, red), bd = block_sum(e_den, red), bb = block_sum((double)nbad, red);
    if (tid == 0) { selbo[0] = bn; selbo[1] = bd; selbo[2] = bn - bd; selbo[3] = bb; }
    // ---------------- phase 3: partial statistics, thread per (component, entry): entry 0 = N_k, 1 = W_k, 2.. = sum r x, sum r x x^T
    for (int e = tid; e < K * SL; e += blockDim.x) {
        const int k = e / SL, j = e - k * SL;
        double acc = 0.0;
        if (j < 2) {
            for (int pl = 0; pl < npts; ++pl) acc += (double)rr[(size_t)pl * K + k];
        } else if (j < 2 + D) {
            for (int pl = 0; pl < npts; ++pl) acc += (double)rr[(size_t)pl * K + k] * (double)xs[(size_t)pl * D + (j - 2)];
        } else {
            const int a = (j - 2 - D) / D, b2 = (j - 2 - D) - a * D;
            for (int pl = 0; pl < npts; ++pl)
                acc += (double)rr[(size_t)pl * K + k] * (double)xs[(size_t)pl * D + a] * (double)xs[(size_t)pl * D + b2];
        }
        sstat[e] = acc;
    }
    cluster.sync();
    // ---------------- phase 4: reduce over the cluster through DSMEM, store, natural-gradient update of this CTA's slice
    const double rho = p.rho_dev != nullptr ? *p.rho_dev : p.rho;
    for (int e = tid + crank * (int)blockDim.x; e < K * SL; e += C * (int)blockDim.x) {
        double part[16];                                           // all remote reads in flight at once (DSMEM latency ~200 cycles each)
#pragma unroll
        for (int c = 0; c < 16; ++c) part[c] = c < C ? cluster.map_shared_rank(sstat, c)[e] : 0.0;
        double tot = 0.0;
#pragma unroll
        for (int c = 0; c < 16; ++c) tot += part[c];
        p.stats[e] = tot;
        const int k = e / SL, j = e - k * SL;
        if (j == 0) {
            p.u0[k] = (T)((1.0 - rho) * (double)p.u0[k] + rho * ((double)p.p0[k] + tot));
            if (!p.only_alpha) {
                p.u3[k] = (T)((1.0 - rho) * (double)p.u3[k] + rho * ((double)p.p3[k] + tot));
                p.u4[k] = (T)((1.0 - rho) * (double)p.u4[k] + rho * ((double)p.p4[k] + tot + 1.0));
            }
        } else if (j >= 2 && !p.only_alpha) {
            if (j < 2 + D) {
                const size_t o = (size_t)k * D + (j - 2);
                p.u2[o] = (T)((1.0 - rho) * (double)p.u2[o] + rho * ((double)p.p2[o] + tot));
            } else {
                const size_t o = (size_t)k * D * D + (j - 2 - D);
                p.u1[o] = (T)((1.0 - rho) * (double)p.u1[o] + rho * ((double)p.p1[o] + tot));
            }
        }
    }
    if (crank == 0 && tid < 4) {
        double tot = 0.0;
        for (int c = 0; c < C; ++c) tot += cluster.map_shared_rank(selbo, c)[tid];
        p.elbo_acc[tid] = tot;
    }
    cluster.sync();                                   // nobody leaves while its shared memory may still be read
}

template <typename T> size_t small_step_smem(int K, int D, int ppc) {
    const int PL = D * D + 2 * D + 4, TL = D * D + D + 4, SL = D * D + D + 2;
    return sizeof(double) * ((size_t)K * SL + 4) + sizeof(T) * ((size_t)K * (PL + TL) + (size_t)ppc * K * (5 + D) + (size_t)ppc * D) + 16;
}

template <typename T, int D>
static int launch_small_step(SmallStepParams<T> p, cudaStream_t st) {
    const int64_t pairs = (int64_t)p.N * p.K;
    constexpr int THREADS = SmThreads<T>::value;
    p.split = p.S < 8 ? p.S : 8;
    while (p.split > 1 && pairs * p.split > (int64_t)16 * THREADS * 2) --p.split;      // at most ~2 work items per thread
    int C = 1;
    while (C < SM_MAX_CLUSTER && pairs * p.split > (int64_t)C * THREADS) C *= 2;       // about one work item per thread
    if ((int64_t)((p.N + C - 1) / C) * p.K > SM_MAX_PAIRS_PER_CTA) return -100;
    p.ppc = (p.N + C - 1) / C;
    const size_t smem = small_step_smem<T>(p.K, D, p.ppc);
    if (smem > 200 * 1024) return -100;
    auto kern = svae_small_step_kernel<T, D>;
    // function attributes are set once per instantiation (and again only if a larger size is needed): at these shapes every
    // extra runtime call is a measurable part of the step
    static size_t smem_set = 0;
    static bool nonportable_set = false;
    if (smem > 48 * 1024 && smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        smem_set = smem;
    }
    if (C > 8 && !nonportable_set) {                                        // 16 CTAs: non-portable cluster size
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return (int)e;
        nonportable_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(C, 1, 1);
    cfg.blockDim = dim3(THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
    return (int)e;                                                          // launch errors are returned by the launch call itself
}

template <typename T>
int svae_small_step(int64_t N, int K, int D, int S, int den_mode, int only_alpha, const T* eta1, const T* eta2d,
                    const T* eta1_phi2, const T* L_raw, const T* pi_raw, const T* const* theta, const T* const* prior,
                    T* const* theta_out, double rho, const double* rho_dev, const T* noise, const T* gum_u, uint64_t seed,
                    int64_t point_offset, T* log_r, T* x_sample, int32_t* z, T* x_k_samples, double* stats, double* elbo_acc,
                    void* stream) {
    if (N <= 0 || K <= 0 || S <= 0 || point_offset < 0) return VMP_E_BADARG;
    if (D < 1 || D > 8 || K > SM_MAX_K || N * (int64_t)K > 8192) return -100;           // caller uses the multi-launch path
    if (den_mode != VMP_DEN_GAUSS && den_mode != VMP_DEN_STUDENT) return VMP_E_BADMODE;
    if (!eta1 || !eta2d || !eta1_phi2 || !L_raw || !pi_raw || !theta || !prior || !theta_out || !log_r || !stats || !elbo_acc)
        return VMP_E_BADARG;
    const int nth = den_mode == VMP_DEN_GAUSS ? 5 : 4;
    for (int i = 0; i < nth; ++i) if (!theta[i]) return VMP_E_BADARG;
    if (!prior[0] || !theta_out[0]) return VMP_E_BADARG;
    if (!only_alpha) for (int i = 0; i < 5; ++i) if (!prior[i] || !theta_out[i]) return VMP_E_BADARG;
    SmallStepParams<T> p;
    p.N = (int)N; p.K = K; p.S = S; p.den_mode = den_mode; p.only_alpha = only_alpha; p.ppc = 0; p.split = 1;
    p.eta1 = eta1; p.eta2d = eta2d; p.eta1_phi2 = eta1_phi2; p.L_raw = L_raw; p.pi_raw = pi_raw;
    p.th0 = theta[0]; p.th1 = theta[1]; p.th2 = theta[2]; p.th3 = theta[3]; p.th4 = nth == 5 ? theta[4] : nullptr;
    p.p0 = prior[0]; p.p1 = only_alpha ? nullptr : prior[1]; p.p2 = only_alpha ? nullptr : prior[2];
    p.p3 = only_alpha ? nullptr : prior[3]; p.p4 = only_alpha ? nullptr : prior[4];
    p.u0 = theta_out[0]; p.u1 = only_alpha ? nullptr : theta_out[1]; p.u2 = only_alpha ? nullptr : theta_out[2];
    p.u3 = only_alpha ? nullptr : theta_out[3]; p.u4 = only_alpha ? nullptr : theta_out[4];
    p.noise = noise; p.gum_u = gum_u; p.seed = seed; p.pair_offset = (uint64_t)point_offset * (uint64_t)K;
    p.rho = rho; p.rho_dev = rho_dev;
    p.log_r = log_r; p.x_sample = x_sample; p.z = z; p.x_k_samples = x_k_samples; p.stats = stats; p.elbo_acc = elbo_acc;
    cudaStream_t st = (cudaStream_t)stream;
#define VMP_SMALL(DD) case DD: return launch_small_step<T, DD>(p, st)
    switch (D) {
        VMP_SMALL(1); VMP_SMALL(2); VMP_SMALL(3); VMP_SMALL(4); VMP_SMALL(5); VMP_SMALL(6); VMP_SMALL(7); VMP_SMALL(8);
        default: return -100;
    }
#undef VMP_SMALL
}

}  // namespace vmp

extern "C" {
int vmp_svae_small_step_supported(int64_t N, int K, int D) {
    return (D >= 1 && D <= 8 && K >= 1 && K <= vmp::SM_MAX_K && N >= 1 && N * (int64_t)K <= 8192) ? 1 : 0;
}
int vmp_svae_small_step_f32(int64_t N, int K, int D, int S, int den_mode, int only_alpha, const float* eta1,
                            const float* eta2_diag, const float* eta1_phi2, const float* L_raw, const float* pi_raw,
                            const float* const* theta, const float* const* prior, float* const* theta_out, double rho,
                            const double* rho_dev, const float* noise, const float* gumbel_u, uint64_t seed,
                            int64_t point_offset, float* log_r, float* x_sample, int32_t* z, float* x_k_samples, double* stats,
                            double* elbo_acc, void* stream) {
    const int rc = vmp::svae_small_step<float>(N, K, D, S, den_mode, only_alpha, eta1, eta2_diag, eta1_phi2, L_raw, pi_raw, theta,
                                               prior, theta_out, rho, rho_dev, noise, gumbel_u, seed, point_offset, log_r, x_sample,
                                               z, x_k_samples, stats, elbo_acc, stream);
    return rc == -100 ? VMP_E_BADARG : rc;
}
int vmp_svae_small_step_f64(int64_t N, int K, int D, int S, int den_mode, int only_alpha, const double* eta1,
                            const double* eta2_diag, const double* eta1_phi2, const double* L_raw, const double* pi_raw,
                            const double* const* theta, const double* const* prior, double* const* theta_out, double rho,
                            const double* rho_dev, const double* noise, const double* gumbel_u, uint64_t seed,
                            int64_t point_offset, double* log_r, double* x_sample, int32_t* z, double* x_k_samples,
                            double* stats, double* elbo_acc, void* stream) {
    const int rc = vmp::svae_small_step<double>(N, K, D, S, den_mode, only_alpha, eta1, eta2_diag, eta1_phi2, L_raw, pi_raw, theta,
                                                prior, theta_out, rho, rho_dev, noise, gumbel_u, seed, point_offset, log_r,
                                                x_sample, z, x_k_samples, stats, elbo_acc, stream);
    return rc == -100 ? VMP_E_BADARG : rc;
}
int vmp_svae_small_step_packed(const VmpSmallStepArgs* a) {
    if (!a) return VMP_E_BADARG;
    if (a->dtype == 0)
        return vmp_svae_small_step_f32(a->N, a->K, a->D, a->S, a->den_mode, a->only_alpha, (const float*)a->eta1,
                                       (const float*)a->eta2_diag, (const float*)a->eta1_phi2, (const float*)a->L_raw,
                                       (const float*)a->pi_raw, (const float* const*)a->theta, (const float* const*)a->prior,
                                       (float* const*)a->theta_out, a->rho, a->rho_dev, (const float*)a->noise,
                                       (const float*)a->gumbel_u, a->seed, a->point_offset, (float*)a->log_r, (float*)a->x_sample,
                                       a->z, (float*)a->x_k_samples, a->stats, a->elbo_acc, a->stream);
    if (a->dtype == 1)
        return vmp_svae_small_step_f64(a->N, a->K, a->D, a->S, a->den_mode, a->only_alpha, (const double*)a->eta1,
                                       (const double*)a->eta2_diag, (const double*)a->eta1_phi2, (const double*)a->L_raw,
                                       (const double*)a->pi_raw, (const double* const*)a->theta, (const double* const*)a->prior,
                                       (double* const*)a->theta_out, a->rho, a->rho_dev, (const double*)a->noise,
                                       (const double*)a->gumbel_u, a->seed, a->point_offset, (double*)a->log_r,
                                       (double*)a->x_sample, a->z, (double*)a->x_k_samples, a->stats, a->elbo_acc, a->stream);
    return VMP_E_BADMODE;
}
}
