// local_step.cu — the fused local VMP step (svae.e_step + subsample_x + ELBO regulariser): entry points, dispatch,
// and the generic thread-per-pair kernels.
//
// Two implementations of the same arithmetic (pair_math.cuh states it):
//   * engine (local_step_fast.cuh): fp32, 9 <= D <= 64: register-resident group engine of dimension 16 / 32 / 64 (the
//           caller's D is embedded with an identity block), TMA-staged records, online Gumbel-max selection — the path
//           bench.py measures.  Taken whenever the caller passes the workspace the staged records need;
//   * generic (this file): fp64, D <= 8, caller-supplied samples (x_in), or no workspace: one thread per
//           (point, component) pair; D <= 8 fully unrolled in registers (the C1-C3 shapes), larger D through local memory.  Kernel A computes scores / samples / ELBO
//           terms and the per-point log-sum-exp; kernel B draws z_n by Gumbel-max from log_r and re-evaluates the
//           selected pair for x[n, z_n, 0].
// The categorical draw follows tf.multinomial's GPU kernel (multinomial_op_gpu.cu.cc): z = argmax_k(logit_k + G_k),
// G_k = -log(-log(u_k)); u[N,K] may be injected for parity tests, otherwise Philox(seed, pair).
#include "local_step_fast.cuh"
#include "pair_math.cuh"

namespace vmp {

constexpr int LS_THREADS = 128;

template <typename T> struct NoiseSrc {
    const T* noise;     // [N,K,D,S] or nullptr
    uint64_t seed;
    int K, D, S;
    int64_t n_offset;   // global index of point 0: the in-kernel stream is keyed by the GLOBAL pair index
    __device__ __forceinline__ uint64_t gpair(int64_t n, int k) const { return (uint64_t)(n + n_offset) * K + k; }
    __device__ __forceinline__ void load(int64_t n, int k, int s, T* eps) const {
        if (noise != nullptr) {
            const T* p = noise + ((uint64_t)n * K + k) * (uint64_t)D * S + s;
            for (int i = 0; i < D; ++i) eps[i] = p[(size_t)i * S];
        } else {
            const uint64_t pair = gpair(n, k);
            for (int q = 0; q < (D + 3) / 4; ++q) {
                const float4 v = philox_normal4(seed, pair, (uint32_t)s, (uint32_t)q);
                const int i = 4 * q;
                eps[i] = (T)v.x;
                if (i + 1 < D) eps[i + 1] = (T)v.y;
                if (i + 2 < D) eps[i + 2] = (T)v.z;
                if (i + 3 < D) eps[i + 3] = (T)v.w;
            }
        }
    }
};

template <typename T, int DT>
__global__ void __launch_bounds__(LS_THREADS)
local_step_kernel(int64_t N, int K, int Drt, int S, int PTS,
                  const T* __restrict__ eta1, const T* __restrict__ eta2d,
                  const T* __restrict__ phi_rec, const T* __restrict__ theta_rec, int den_mode,
                  NoiseSrc<T> nz, const T* __restrict__ x_in, T* __restrict__ log_r, T* __restrict__ x_k_samples,
                  double* __restrict__ elbo_acc) {
    using PM = PairMath<T, DT>;
    const int D = DT ? DT : Drt;
    extern __shared__ unsigned char smraw[];
    T* sc = reinterpret_cast<T*>(smraw);         // [PTS*K] score -> log r
    T* tnum = sc + (size_t)PTS * K;              // [PTS*K] mean_s log N(x|phi~) (without log r)
    T* tden = tnum + (size_t)PTS * K;            // [PTS*K] mean_s log p(x|theta_k) + E log pi_k
    __shared__ double red[32];

    const int64_t pt0 = (int64_t)blockIdx.x * PTS;
    const int npts = (int)min((int64_t)PTS, N - pt0);
    const int npairs = npts * K;
    const int plen = phi_record_len(D), tlen = theta_record_len(D);
    int nbad = 0;

    for (int p = threadIdx.x; p < npairs; p += blockDim.x) {
        const int pl = p / K, k = p - pl * K;
        const int64_t n = pt0 + pl;
        PM pm;
        pm.factor(D, eta1 + n * D, eta2d + n * D, phi_rec + (size_t)k * plen);
        nbad += pm.bad;
        const T* trec = theta_rec + (size_t)k * tlen;
        const T cden = trec[D * D + D], nu = trec[D * D + D + 1];
        T eps[PM::DM], x[PM::DM];
        T snum = T(0), sden = T(0);
        for (int s = 0; s < S; ++s) {
            if (x_in != nullptr) {
                const T* xi = x_in + (((size_t)n * K + k) * S + s) * D;
#pragma unroll
                for (int i = 0; i < D; ++i) x[i] = xi[i];
                pm.eps_from_x(D, x, eps);
            } else {
                nz.load(n, k, s, eps);
                pm.sample(D, eps, x);
            }
            T e2 = T(0);
#pragma unroll
            for (int i = 0; i < D; ++i) e2 = fma(eps[i], eps[i], e2);
            if (x_k_samples != nullptr) {
                T* xo = x_k_samples + (((size_t)n * K + k) * S + s) * D;
#pragma unroll
                for (int i = 0; i < D; ++i) xo[i] = x[i];
            }
            snum += T(-0.5) * e2;
            sden += den_logprob<T>(den_mode, D, PM::maha(D, trec, x), cden, nu);
        }
        sc[p] = pm.score;
        tnum[p] = snum / T(S) + pm.hld - T(0.5 * VMP_LOG_2PI) * T(D);
        tden[p] = sden / T(S);
    }
    __syncthreads();

    // per-point log-sum-exp (gaussian.py:67-71), sums in double
    for (int pl = threadIdx.x; pl < npts; pl += blockDim.x) {
        T* s = sc + (size_t)pl * K;
        T mx = s[0];
        for (int k = 1; k < K; ++k) mx = max(mx, s[k]);
        double se = 0.0;
        for (int k = 0; k < K; ++k) se += (double)t_exp(s[k] - mx);     // T-precision exp, double accumulation
        const T lse = mx + (T)log(se);
        for (int k = 0; k < K; ++k) s[k] -= lse;
    }
    __syncthreads();

    double e_num = 0.0, e_den = 0.0;
    for (int p = threadIdx.x; p < npairs; p += blockDim.x) {
        const T lr = sc[p];
        log_r[pt0 * K + p] = lr;
        const double r = (double)t_exp(lr);
        e_num += r * ((double)tnum[p] + (double)lr);
        e_den += r * (double)tden[p];
    }
    const double bn = block_sum(e_num, red);
    const double bd = block_sum(e_den, red);
    const double bb = block_sum((double)nbad, red);
    if (threadIdx.x == 0) {
        atomicAdd(elbo_acc + 0, bn);
        atomicAdd(elbo_acc + 1, bd);
        atomicAdd(elbo_acc + 2, bn - bd);
        if (bb != 0.0) atomicAdd(elbo_acc + 3, bb);
    }
}

template <typename T, int DT>
__global__ void __launch_bounds__(LS_THREADS)
select_sample_kernel(int64_t N, int K, int Drt, int S, const T* __restrict__ eta1, const T* __restrict__ eta2d,
                     const T* __restrict__ phi_rec, const T* __restrict__ log_r, const T* __restrict__ gum_u,
                     NoiseSrc<T> nz, T* __restrict__ x_sample, int32_t* __restrict__ z_out) {
    using PM = PairMath<T, DT>;
    const int D = DT ? DT : Drt;
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    // Gumbel-max: first arg-max of log_r + G (normalisation does not move the arg-max)
    int z = 0;
    T best = -CUDART_INF_F;
    for (int k = 0; k < K; ++k) {
        const uint64_t pair = (uint64_t)n * K + k;
        const T u = gum_u != nullptr ? gum_u[pair] : (T)philox_uniform_pair(nz.seed, nz.gpair(n, k));
        const T cand = log_r[pair] + gumbel_from_uniform<T>(u);
        if (cand > best) { best = cand; z = k; }
    }
    if (z_out != nullptr) z_out[n] = z;
    if (x_sample == nullptr) return;
    PM pm;
    pm.factor(D, eta1 + n * D, eta2d + n * D, phi_rec + (size_t)z * phi_record_len(D));
    T eps[PM::DM], x[PM::DM];
    nz.load(n, z, 0, eps);
    pm.sample(D, eps, x);
#pragma unroll
    for (int i = 0; i < D; ++i) x_sample[n * D + i] = x[i];
}

template <typename T>
__global__ void fill_noise_kernel(int64_t N, int K, int D, int S, uint64_t seed, int64_t n_offset, T* noise, T* u) {
    const uint64_t poff = (uint64_t)n_offset * K;
    const int64_t total = N * K * (int64_t)D * S;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (noise != nullptr)
        for (int64_t e = t0; e < total; e += stride) {
            const int s = (int)(e % S);
            const int d = (int)((e / S) % D);
            const int64_t pair = e / ((int64_t)S * D);
            noise[e] = (T)philox_normal1(seed, (uint64_t)pair + poff, (uint32_t)s, (uint32_t)d);
        }
    if (u != nullptr)
        for (int64_t pr = t0; pr < N * K; pr += stride) u[pr] = (T)philox_uniform_pair(seed, (uint64_t)pr + poff);
}

template <typename T, int DT>
static int launch_generic(int64_t N, int K, int D, int S, const T* eta1, const T* eta2d, const T* phi_rec,
                          const T* theta_rec, int den_mode, const T* noise, const T* gum_u, uint64_t seed,
                          int64_t n_offset, const T* x_in, T* log_r, T* x_sample, int32_t* z, T* x_k_samples,
                          double* elbo_acc, cudaStream_t st) {
    NoiseSrc<T> nz{noise, seed, K, D, S, n_offset};
    int PTS = LS_THREADS / K;
    if (PTS < 1) PTS = 1;
    const size_t smem = (size_t)3 * PTS * K * sizeof(T);
    if (smem > 200 * 1024) return VMP_E_BADARG;
    auto kern = local_step_kernel<T, DT>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    const int64_t grid = (N + PTS - 1) / PTS;
    if (grid > 0x7fffffffLL) return VMP_E_BADARG;
    kern<<<(unsigned)grid, LS_THREADS, smem, st>>>(N, K, D, S, PTS, eta1, eta2d, phi_rec, theta_rec, den_mode, nz,
                                                   x_in, log_r, x_k_samples, elbo_acc);
    if (int e = launch_status()) return e;
    if (x_sample != nullptr || z != nullptr) {
        const int64_t g2 = (N + LS_THREADS - 1) / LS_THREADS;
        select_sample_kernel<T, DT><<<(unsigned)g2, LS_THREADS, 0, st>>>(N, K, D, S, eta1, eta2d, phi_rec, log_r,
                                                                       gum_u, nz, x_sample, z);
        if (int e = launch_status()) return e;
    }
    return VMP_OK;
}

// engine dimension for a caller dimension: 0 = no engine (D <= 8 runs the register-unrolled generic kernels)
static int engine_dim(int D) { return D <= 8 ? 0 : D <= 16 ? 16 : D <= 32 ? 32 : 64; }

static size_t fast_workspace_bytes(int K, int D) {
    const int De = engine_dim(D);
    return De ? sizeof(float) * (size_t)K * fast_rec_len(De) : 0;
}

// the group engine when the call qualifies (fp32, D > 8, drawn samples, workspace given); -100 when it does not
static int try_fast(int64_t N, int K, int D, int S, const float* eta1, const float* eta2d, const float* phi_rec,
                    const float* theta_rec, int den_mode, const float* noise, const float* gum_u, uint64_t seed,
                    int64_t n_offset, const float* x_in, float* log_r, float* x_sample, int32_t* z, float* x_k_samples,
                    double* elbo_acc, void* work, size_t work_bytes, cudaStream_t st) {
    if (x_in != nullptr || work == nullptr) return -100;
    const int De = engine_dim(D);
    const size_t need = fast_workspace_bytes(K, D);
    if (De == 0 || work_bytes < need) return -100;
    const size_t smem = De == 64 ? fast_smem_bytes<64, 16>(K) : De == 32 ? fast_smem_bytes<32, VMP_D32_LANES>(K) : fast_smem_bytes<16, 4>(K);
    const int minb = De == 64 ? 1 : (De == 32 ? FastLaunch<32, VMP_D32_LANES>::MINB : 2);
    if (smem * minb > 220 * 1024) return -100;      // very large K: the per-point score table no longer fits
    float* recs = static_cast<float*>(work);
    launch_pack_fast_records(K, D, De, phi_rec, theta_rec, recs, st);
    if (int e = launch_status()) return e;
    FastParams p{N, K, S, den_mode, D, (uint64_t)n_offset * (uint64_t)K, eta1, eta2d, recs, noise, gum_u, seed, log_r, x_sample, z, x_k_samples, elbo_acc, 0};
    if (De == 64) return launch_fast<64, 16>(p, st);
    if (De == 32) return launch_fast<32, VMP_D32_LANES>(p, st);
    return launch_fast<16, 4>(p, st);
}
template <typename T>
static int try_fast_t(int64_t, int, int, int, const T*, const T*, const T*, const T*, int, const T*, const T*, uint64_t,
                      int64_t, const T*, T*, T*, int32_t*, T*, double*, void*, size_t, cudaStream_t) {
    return -100;
}
template <>
int try_fast_t<float>(int64_t N, int K, int D, int S, const float* a, const float* b, const float* c, const float* d,
                      int m, const float* e, const float* f, uint64_t seed, int64_t n_offset, const float* g, float* h,
                      float* i, int32_t* z, float* j, double* acc, void* work, size_t wb, cudaStream_t st) {
    return try_fast(N, K, D, S, a, b, c, d, m, e, f, seed, n_offset, g, h, i, z, j, acc, work, wb, st);
}

template <typename T>
int svae_local_step(int64_t N, int K, int D, int S, const T* eta1, const T* eta2d, const T* phi_rec,
                    const T* theta_rec, int den_mode, const T* noise, const T* gum_u, uint64_t seed, int64_t n_offset,
                    const T* x_in, T* log_r, T* x_sample, int32_t* z, T* x_k_samples, double* elbo_acc, void* work,
                    size_t work_bytes, void* stream) {
    if (N < 0 || K <= 0 || S <= 0 || n_offset < 0) return VMP_E_BADARG;
    if (D < 1 || D > VMP_MAX_D) return VMP_E_BADDIM;
    if (den_mode != VMP_DEN_GAUSS && den_mode != VMP_DEN_STUDENT) return VMP_E_BADMODE;
    if (N == 0) return VMP_OK;
    if (!eta1 || !eta2d || !phi_rec || !theta_rec || !log_r || !elbo_acc) return VMP_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int rc = try_fast_t<T>(N, K, D, S, eta1, eta2d, phi_rec, theta_rec, den_mode, noise, gum_u, seed, n_offset, x_in,
                                 log_r, x_sample, z, x_k_samples, elbo_acc, work, work_bytes, st);
    if (rc != -100) return rc;
#define VMP_LS(DD)                                                                                              \
    case DD:                                                                                                    \
        return launch_generic<T, DD>(N, K, D, S, eta1, eta2d, phi_rec, theta_rec, den_mode, noise, gum_u, seed, \
                                     n_offset, x_in, log_r, x_sample, z, x_k_samples, elbo_acc, st)
    switch (D) {
        VMP_LS(1); VMP_LS(2); VMP_LS(3); VMP_LS(4); VMP_LS(5); VMP_LS(6); VMP_LS(7); VMP_LS(8);
        default:
            return launch_generic<T, 0>(N, K, D, S, eta1, eta2d, phi_rec, theta_rec, den_mode, noise, gum_u, seed,
                                        n_offset, x_in, log_r, x_sample, z, x_k_samples, elbo_acc, st);
    }
#undef VMP_LS
}

template <typename T>
int fill_noise(int64_t N, int K, int D, int S, uint64_t seed, int64_t n_offset, T* noise, T* u, void* stream) {
    if (N <= 0 || K <= 0 || D <= 0 || S <= 0 || n_offset < 0 || (!noise && !u)) return VMP_E_BADARG;
    fill_noise_kernel<T><<<148 * 8, 256, 0, (cudaStream_t)stream>>>(N, K, D, S, seed, n_offset, noise, u);
    return launch_status();
}

}  // namespace vmp

extern "C" {
size_t vmp_svae_local_step_workspace_bytes(int K, int D) {
    const size_t b = vmp::fast_workspace_bytes(K, D);
    return b ? b : 16;
}
int vmp_svae_local_step_f32(int64_t N, int K, int D, int S, const float* eta1, const float* eta2_diag,
                            const float* phi_rec, const float* theta_rec, int den_mode, const float* noise,
                            const float* gumbel_u, uint64_t seed, int64_t point_offset, const float* x_in, float* log_r,
                            float* x_sample, int32_t* z, float* x_k_samples, double* elbo_acc, void* workspace,
                            size_t workspace_bytes, void* stream) {
    return vmp::svae_local_step<float>(N, K, D, S, eta1, eta2_diag, phi_rec, theta_rec, den_mode, noise, gumbel_u, seed,
                                       point_offset, x_in, log_r, x_sample, z, x_k_samples, elbo_acc, workspace, workspace_bytes,
                                       stream);
}
int vmp_svae_local_step_f64(int64_t N, int K, int D, int S, const double* eta1, const double* eta2_diag,
                            const double* phi_rec, const double* theta_rec, int den_mode, const double* noise,
                            const double* gumbel_u, uint64_t seed, int64_t point_offset, const double* x_in, double* log_r,
                            double* x_sample, int32_t* z, double* x_k_samples, double* elbo_acc, void* workspace,
                            size_t workspace_bytes, void* stream) {
    return vmp::svae_local_step<double>(N, K, D, S, eta1, eta2_diag, phi_rec, theta_rec, den_mode, noise, gumbel_u,
                                        seed, point_offset, x_in, log_r, x_sample, z, x_k_samples, elbo_acc, workspace,
                                        workspace_bytes, stream);
}
int vmp_fill_noise_f32(int64_t N, int K, int D, int S, uint64_t seed, int64_t point_offset, float* noise, float* u,
                       void* stream) {
    return vmp::fill_noise<float>(N, K, D, S, seed, point_offset, noise, u, stream);
}
int vmp_fill_noise_f64(int64_t N, int K, int D, int S, uint64_t seed, int64_t point_offset, double* noise, double* u,
                       void* stream) {
    return vmp::fill_noise<double>(N, K, D, S, seed, point_offset, noise, u, stream);
}
}
