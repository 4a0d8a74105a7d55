// prepare.cu — K-sized prologues: phi_gmm / theta -> per-component records, batched SPD inverse.
// One CTA per component; all arithmetic in double in shared memory (K*D^3 work, not perf relevant),
// results rounded once to the storage type T.
#include "block_linalg.cuh"
#include "common.cuh"

namespace vmp {

constexpr int PREP_THREADS = 128;

// tril(L_raw) with softplus on the diagonal (svae.py:349-350 / 369-370), into smem (double)
template <typename T>
__device__ void load_tril_softplus(const T* L_raw, double* Ls, int D, int ld) {
    for (int e = threadIdx.x; e < D * D; e += blockDim.x) {
        const int i = e / D, j = e % D;
        double v = 0.0;
        if (j < i) v = (double)L_raw[e];
        else if (j == i) v = t_softplus<double>((double)L_raw[e]);
        Ls[i * ld + j] = v;
    }
    __syncthreads();
}

__device__ double block_digamma_term(const double alpha_k, double sum_alpha) {
    return digamma_pos(alpha_k) - digamma_pos(sum_alpha);
}

template <typename T>
__global__ void __launch_bounds__(PREP_THREADS)
phi_prepare_kernel(int K, int D, const T* __restrict__ eta1_phi2, const T* __restrict__ L_raw,
                   const T* __restrict__ pi_raw, T* __restrict__ rec) {
    extern __shared__ double sm[];
    const int ld = D + 1;
    double* Ls = sm;                 // D x ld
    double* vec = Ls + D * ld;       // D
    double* red = vec + D;           // 32
    const int k = blockIdx.x;
    T* out = rec + (size_t)k * phi_record_len(D);
    load_tril_softplus(L_raw + (size_t)k * D * D, Ls, D, ld);
    // P2 = L L^T
    for (int e = threadIdx.x; e < D * D; e += blockDim.x) {
        const int i = e / D, j = e % D;
        const int m = i < j ? i : j;
        double s = 0.0;
        for (int c = 0; c <= m; ++c) s += Ls[i * ld + c] * Ls[j * ld + c];
        out[e] = (T)s;
    }
    // mu2 = P2^-1 h2 = L^-T (L^-1 h2); logdet P2 = 2 sum log L_ii   (thread 0; D^2 work)
    if (threadIdx.x == 0) {
        double ld2 = 0.0;
        for (int i = 0; i < D; ++i) {
            double s = (double)eta1_phi2[(size_t)k * D + i];
            for (int c = 0; c < i; ++c) s -= Ls[i * ld + c] * vec[c];
            vec[i] = s / Ls[i * ld + i];
            ld2 += log(Ls[i * ld + i]);
        }
        for (int i = D - 1; i >= 0; --i) {
            double s = vec[i];
            for (int c = i + 1; c < D; ++c) s -= Ls[c * ld + i] * vec[c];
            vec[i] = s / Ls[i * ld + i];
        }
        out[D * D + 2 * D + 1] = (T)(2.0 * ld2);
        out[D * D + 2 * D + 2] = (T)0;
        out[D * D + 2 * D + 3] = (T)0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        out[D * D + i] = (T)vec[i];
        out[D * D + D + i] = eta1_phi2[(size_t)k * D + i];
    }
    // log softmax(pi_raw)[k]  (svae.py:356 + gaussian.py:64 take log(softmax))
    double mx = -CUDART_INF;
    for (int j = threadIdx.x; j < K; j += blockDim.x) mx = fmax(mx, (double)pi_raw[j]);
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmax(mx, red[w]);
    __syncthreads();
    double se = 0.0;
    for (int j = threadIdx.x; j < K; j += blockDim.x) se += exp((double)pi_raw[j] - mx);
    se = block_sum(se, red);
    if (threadIdx.x == 0) out[D * D + 2 * D] = (T)((double)pi_raw[k] - mx - log(se));
}

// sum over K of (alpha_nat + 1), every CTA recomputes it (K small)
template <typename T>
__device__ double sum_alpha_std(const T* alpha_nat, int K, double* red) {
    double s = 0.0;
    for (int j = threadIdx.x; j < K; j += blockDim.x) s += (double)alpha_nat[j] + 1.0;
    s = block_sum(s, red);
    __shared__ double bc;
    if (threadIdx.x == 0) bc = s;
    __syncthreads();
    return bc;
}

template <typename T>
__global__ void __launch_bounds__(PREP_THREADS)
theta_prepare_gauss_kernel(int K, int D, const T* __restrict__ alpha, const T* __restrict__ A,
                           const T* __restrict__ b, const T* __restrict__ beta, const T* __restrict__ v_hat,
                           T* __restrict__ rec) {
    extern __shared__ double sm[];
    const int ld = D + 1;
    double* C = sm;                  // D x ld   -> chol(C)
    double* W = C + D * ld;          // D x ld
    double* red = W + D * ld;        // 32
    const int k = blockIdx.x;
    T* out = rec + (size_t)k * theta_record_len(D);
    const double bk = (double)beta[k];
    const double v = (double)v_hat[k] - D - 2.0;                     // niw.py:42
    // C = A - outer(b, m), m = b / beta                               niw.py:35-41
    for (int e = threadIdx.x; e < D * D; e += blockDim.x) {
        const int i = e / D, j = e % D;
        const double bi = (double)b[(size_t)k * D + i], bj = (double)b[(size_t)k * D + j];
        C[i * ld + j] = (double)A[(size_t)k * D * D + e] - bi * (bj / bk);
    }
    __syncthreads();
    chol_lower_block(C, D, ld);
    tri_inverse_block(C, W, D, ld);
    // E[Sigma] = C / v (niw.py:8-17)  =>  P_theta = v C^-1 = (sqrt(v) Lc^-1)^T (sqrt(v) Lc^-1)
    const double sv = sqrt(v);
    for (int e = threadIdx.x; e < D * D; e += blockDim.x) {
        const int i = e / D, j = e % D;
        out[e] = (T)(j <= i ? sv * W[i * ld + j] : 0.0);
    }
    for (int i = threadIdx.x; i < D; i += blockDim.x) out[D * D + i] = (T)((double)b[(size_t)k * D + i] / bk);
    const double sa = sum_alpha_std(alpha, K, red);
    if (threadIdx.x == 0) {
        double hl = 0.0;
        for (int i = 0; i < D; ++i) hl += log(C[i * ld + i]);
        const double logdetP = D * log(v) - 2.0 * hl;
        const double elogpi = digamma_pos((double)alpha[k] + 1.0) - digamma_pos(sa);   // dirichlet.py:8-12
        out[D * D + D + 0] = (T)(0.5 * logdetP - 0.5 * D * VMP_LOG_2PI + elogpi);
        out[D * D + D + 1] = (T)0;
        out[D * D + D + 2] = (T)elogpi;
        out[D * D + D + 3] = (T)logdetP;
    }
}

template <typename T>
__global__ void __launch_bounds__(PREP_THREADS)
theta_prepare_student_kernel(int K, int D, const T* __restrict__ alpha, const T* __restrict__ mu,
                             const T* __restrict__ L_raw, const T* __restrict__ dof, T* __restrict__ rec) {
    extern __shared__ double sm[];
    const int ld = D + 1;
    double* Ls = sm;
    double* W = Ls + D * ld;
    double* red = W + D * ld;
    const int k = blockIdx.x;
    T* out = rec + (size_t)k * theta_record_len(D);
    load_tril_softplus(L_raw + (size_t)k * D * D, Ls, D, ld);      // Sigma = L L^T (svae.py:365-371)
    tri_inverse_block(Ls, W, D, ld);                                // Delta^2 = |L^-1 (x - mu)|^2
    for (int e = threadIdx.x; e < D * D; e += blockDim.x) {
        const int i = e / D, j = e % D;
        out[e] = (T)(j <= i ? W[i * ld + j] : 0.0);
    }
    for (int i = threadIdx.x; i < D; i += blockDim.x) out[D * D + i] = mu[(size_t)k * D + i];
    const double sa = sum_alpha_std(alpha, K, red);
    if (threadIdx.x == 0) {
        double hl = 0.0;
        for (int i = 0; i < D; ++i) hl += log(Ls[i * ld + i]);
        const double nu = (double)dof[k];
        const double elogpi = digamma_pos((double)alpha[k] + 1.0) - digamma_pos(sa);
        // student_t.py:34-36 : lgamma((v+D)/2) - lgamma(v/2) - D/2 log(pi v) - 1/2 logdet(Sigma)
        const double c = lgamma(0.5 * (nu + D)) - lgamma(0.5 * nu) - 0.5 * D * (VMP_LOG_PI + log(nu)) - hl + elogpi;
        out[D * D + D + 0] = (T)c;
        out[D * D + D + 1] = (T)nu;
        out[D * D + D + 2] = (T)elogpi;
        out[D * D + D + 3] = (T)(-2.0 * hl);
    }
}

template <typename T>
__global__ void __launch_bounds__(PREP_THREADS)
spd_inverse_kernel(int D, const T* __restrict__ in, T* __restrict__ inv, T* __restrict__ logdet) {
    extern __shared__ double sm[];
    const int ld = D + 1;
    double* C = sm;
    double* W = C + D * ld;
    const int k = blockIdx.x;
    for (int e = threadIdx.x; e < D * D; e += blockDim.x) C[(e / D) * ld + e % D] = (double)in[(size_t)k * D * D + e];
    __syncthreads();
    chol_lower_block(C, D, ld);
    if (logdet != nullptr && threadIdx.x == 0) {
        double hl = 0.0;
        for (int i = 0; i < D; ++i) hl += log(C[i * ld + i]);
        logdet[k] = (T)(2.0 * hl);
    }
    if (inv == nullptr) return;
    tri_inverse_block(C, W, D, ld);
    for (int e = threadIdx.x; e < D * D; e += blockDim.x) {         // inv = W^T W
        const int i = e / D, j = e % D;
        const int m = i > j ? i : j;
        double s = 0.0;
        for (int c = m; c < D; ++c) s += W[c * ld + i] * W[c * ld + j];
        inv[(size_t)k * D * D + e] = (T)s;
    }
}

// Dense-natural-parameter sampling (svae.sample_x_per_comp, svae.py:95-119): one CTA per (n,k) system.
// P = -2 eta2 = L L^T; x_s = P^-1 eta1 + L^-T eps_s = L^-T (L^-1 eta1 + eps_s); thread s handles sample s.
template <typename T>
__global__ void __launch_bounds__(PREP_THREADS)
gaussian_sample_nat_kernel(int D, int S, const T* __restrict__ eta1, const T* __restrict__ eta2,
                           const T* __restrict__ noise, T* __restrict__ x, int* __restrict__ bad) {
    extern __shared__ double sm[];
    const int ld = D + 1;
    double* C = sm;                     // [D][ld]
    double* y = C + D * ld;             // [D]  L^-1 eta1
    double* w = y + D;                  // [PREP_THREADS][D] per-thread right-hand side
    const size_t b = blockIdx.x;
    for (int e = threadIdx.x; e < D * D; e += blockDim.x) C[(e / D) * ld + e % D] = -2.0 * (double)eta2[b * D * D + e];
    __syncthreads();
    chol_lower_block(C, D, ld);
    if (threadIdx.x == 0) {
        bool nonpd = false;
        for (int i = 0; i < D; ++i) {
            double s = (double)eta1[b * D + i];
            for (int c = 0; c < i; ++c) s -= C[i * ld + c] * y[c];
            y[i] = s / C[i * ld + i];
            nonpd |= !(C[i * ld + i] > 0.0);
        }
        if (nonpd && bad != nullptr) atomicAdd(bad, 1);
    }
    __syncthreads();
    for (int s0 = 0; s0 < S; s0 += blockDim.x) {
        const int s = s0 + threadIdx.x;
        if (s < S) {
            double* v = w + (size_t)threadIdx.x * D;
            for (int i = 0; i < D; ++i) v[i] = y[i] + (double)noise[(b * D + i) * S + s];      // noise[N,K,D,S]
            for (int i = D - 1; i >= 0; --i) {                                                  // L^T x = v
                double t = v[i];
                for (int c = i + 1; c < D; ++c) t -= C[c * ld + i] * v[c];
                v[i] = t / C[i * ld + i];
            }
            for (int i = 0; i < D; ++i) x[(b * S + s) * D + i] = (T)v[i];                        // x[N,K,S,D]
        }
    }
}

static size_t prep_smem(int D, int nmat) { return sizeof(double) * ((size_t)nmat * D * (D + 1) + D + 64); }

template <typename K>
static int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}

template <typename T>
int phi_prepare(int K, int D, const T* eta1, const T* L_raw, const T* pi_raw, T* rec, void* stream) {
    if (K <= 0 || !eta1 || !L_raw || !pi_raw || !rec) return VMP_E_BADARG;
    if (D < 1 || D > VMP_MAX_D) return VMP_E_BADDIM;
    const size_t sm = prep_smem(D, 1);
    if (int e = set_smem(phi_prepare_kernel<T>, sm)) return e;
    phi_prepare_kernel<T><<<K, PREP_THREADS, sm, (cudaStream_t)stream>>>(K, D, eta1, L_raw, pi_raw, rec);
    return launch_status();
}

template <typename T>
int theta_prepare_gauss(int K, int D, const T* alpha, const T* A, const T* b, const T* beta, const T* v_hat,
                        T* rec, void* stream) {
    if (K <= 0 || !alpha || !A || !b || !beta || !v_hat || !rec) return VMP_E_BADARG;
    if (D < 1 || D > VMP_MAX_D) return VMP_E_BADDIM;
    const size_t sm = prep_smem(D, 2);
    if (int e = set_smem(theta_prepare_gauss_kernel<T>, sm)) return e;
    theta_prepare_gauss_kernel<T><<<K, PREP_THREADS, sm, (cudaStream_t)stream>>>(K, D, alpha, A, b, beta, v_hat, rec);
    return launch_status();
}

template <typename T>
int theta_prepare_student(int K, int D, const T* alpha, const T* mu, const T* L_raw, const T* dof, T* rec,
                          void* stream) {
    if (K <= 0 || !alpha || !mu || !L_raw || !dof || !rec) return VMP_E_BADARG;
    if (D < 1 || D > VMP_MAX_D) return VMP_E_BADDIM;
    const size_t sm = prep_smem(D, 2);
    if (int e = set_smem(theta_prepare_student_kernel<T>, sm)) return e;
    theta_prepare_student_kernel<T><<<K, PREP_THREADS, sm, (cudaStream_t)stream>>>(K, D, alpha, mu, L_raw, dof, rec);
    return launch_status();
}

template <typename T>
int spd_inverse(int K, int D, const T* in, T* inv, T* logdet, void* stream) {
    if (K <= 0 || !in || (!inv && !logdet)) return VMP_E_BADARG;
    if (D < 1 || D > VMP_MAX_D) return VMP_E_BADDIM;
    const size_t sm = prep_smem(D, 2);
    if (int e = set_smem(spd_inverse_kernel<T>, sm)) return e;
    spd_inverse_kernel<T><<<K, PREP_THREADS, sm, (cudaStream_t)stream>>>(D, in, inv, logdet);
    return launch_status();
}

template <typename T>
int gaussian_sample_nat(int64_t B, int D, int S, const T* eta1, const T* eta2, const T* noise, T* x, int* bad,
                        void* stream) {
    if (B < 0 || S <= 0) return VMP_E_BADARG;
    if (D < 1 || D > VMP_MAX_D) return VMP_E_BADDIM;
    if (B == 0) return VMP_OK;
    if (!eta1 || !eta2 || !noise || !x || B > 0x7fffffffLL) return VMP_E_BADARG;
    const size_t sm = sizeof(double) * ((size_t)D * (D + 1) + D + (size_t)PREP_THREADS * D);
    if (int e = set_smem(gaussian_sample_nat_kernel<T>, sm)) return e;
    gaussian_sample_nat_kernel<T><<<(unsigned)B, PREP_THREADS, sm, (cudaStream_t)stream>>>(D, S, eta1, eta2, noise, x, bad);
    return launch_status();
}

}  // namespace vmp

extern "C" {
int vmp_version(void) { return 200; }
int vmp_gaussian_sample_nat_f32(int64_t B, int D, int S, const float* eta1, const float* eta2, const float* noise,
                                float* x, int* non_pd, void* s) {
    return vmp::gaussian_sample_nat<float>(B, D, S, eta1, eta2, noise, x, non_pd, s);
}
int vmp_gaussian_sample_nat_f64(int64_t B, int D, int S, const double* eta1, const double* eta2, const double* noise,
                                double* x, int* non_pd, void* s) {
    return vmp::gaussian_sample_nat<double>(B, D, S, eta1, eta2, noise, x, non_pd, s);
}
int vmp_phi_record_len(int D) { return vmp::phi_record_len(D); }
int vmp_theta_record_len(int D) { return vmp::theta_record_len(D); }
int vmp_stats_len(int D) { return vmp::stats_len(D); }

int vmp_phi_prepare_f32(int K, int D, const float* e, const float* L, const float* p, float* rec, void* s) {
    return vmp::phi_prepare<float>(K, D, e, L, p, rec, s);
}
int vmp_phi_prepare_f64(int K, int D, const double* e, const double* L, const double* p, double* rec, void* s) {
    return vmp::phi_prepare<double>(K, D, e, L, p, rec, s);
}
int vmp_theta_prepare_gauss_f32(int K, int D, const float* alpha, const float* A, const float* b, const float* beta,
                                const float* v_hat, float* rec, void* s) {
    return vmp::theta_prepare_gauss<float>(K, D, alpha, A, b, beta, v_hat, rec, s);
}
int vmp_theta_prepare_gauss_f64(int K, int D, const double* alpha, const double* A, const double* b,
                                const double* beta, const double* v_hat, double* rec, void* s) {
    return vmp::theta_prepare_gauss<double>(K, D, alpha, A, b, beta, v_hat, rec, s);
}
int vmp_theta_prepare_student_f32(int K, int D, const float* alpha, const float* mu, const float* L_raw,
                                  const float* dof, float* rec, void* s) {
    return vmp::theta_prepare_student<float>(K, D, alpha, mu, L_raw, dof, rec, s);
}
int vmp_theta_prepare_student_f64(int K, int D, const double* alpha, const double* mu, const double* L_raw,
                                  const double* dof, double* rec, void* s) {
    return vmp::theta_prepare_student<double>(K, D, alpha, mu, L_raw, dof, rec, s);
}
int vmp_spd_inverse_f32(int K, int D, const float* in, float* inv, float* logdet, void* s) {
    return vmp::spd_inverse<float>(K, D, in, inv, logdet, s);
}
int vmp_spd_inverse_f64(int K, int D, const double* in, double* inv, double* logdet, void* s) {
    return vmp::spd_inverse<double>(K, D, in, inv, logdet, s);
}
}
