// elbo_terms.cu — ELBO terms outside the fused local step:
//   * decoder-side weighted reductions  vae.expected_diagonal_gaussian_loglike (vae.py:201-250, weighted branch) and
//     vae.expected_bernoulli_loglike (vae.py:175-198): HBM-bound streaming reductions over [N,K,S,Dobs];
//   * the general-form Gaussian log-densities of distributions/gaussian.py (30-71 normalised over K, 74-105 per
//     sample) for arbitrary dense natural parameters eta2[N,K,D,D] — API surface; the SVAE hot path never
//     materialises eta2[N,K,D,D] and uses local_step.cu instead.
#include "common.cuh"

namespace vmp {

// mode 0: sum_{n,k,s,d} w_nk [ (y_nd - mean)^2 / var + log(var + 1e-8) ]      (vae.py:240)
// mode 1: sum_{n,k,s,d} w_nk [ -log(1 + exp(-logit * y_nd)) ]                 (vae.py:191-196, before the 1/S)
template <typename T>
__global__ void __launch_bounds__(256)
decoder_loglike_kernel(int64_t N, int K, int S, int Dobs, int mode, const T* __restrict__ y,
                       const T* __restrict__ means, const T* __restrict__ out2, const T* __restrict__ w,
                       double* __restrict__ acc) {
    __shared__ double red[32];
    const int64_t rows = N * K * (int64_t)S;            // one row = Dobs contiguous decoder outputs
    double local = 0.0;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp; row < rows; row += nwarps) {
        const int64_t nk = row / S;
        const int64_t n = nk / K;
        const T wk = w[nk];
        const T* yr = y + n * Dobs;
        const T* o2 = out2 + row * Dobs;
        T s = T(0);
        if (mode == 0) {
            const T* mu = means + row * Dobs;
            for (int d = lane; d < Dobs; d += 32) {
                const T e = yr[d] - mu[d], v = o2[d];
                s += e * e / v + t_log(v + T(1e-8));
            }
        } else {
            for (int d = lane; d < Dobs; d += 32) s -= t_softplus(-o2[d] * yr[d]);
        }
        local += (double)(wk * s);
    }
    const double b = block_sum(local, red);
    if (threadIdx.x == 0) atomicAdd(acc, b);
}

template <typename T>
int decoder_loglike(int64_t N, int K, int S, int Dobs, int mode, const T* y, const T* means, const T* out2, const T* w,
                    double* acc, void* stream) {
    if (N < 0 || K <= 0 || S <= 0 || Dobs <= 0 || !y || !out2 || !w || !acc) return VMP_E_BADARG;
    if (mode != 0 && mode != 1) return VMP_E_BADMODE;
    if (mode == 0 && !means) return VMP_E_BADARG;
    if (N == 0) return VMP_OK;
    const int64_t rows = N * K * (int64_t)S;
    int64_t grid = (rows + 7) / 8;
    if (grid > 148 * 16) grid = 148 * 16;
    decoder_loglike_kernel<T><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(N, K, S, Dobs, mode, y, means, out2, w, acc);
    return launch_status();
}

// Reverse of the weighted decoder reduction: with F = sum_{n,k,s,d} w_nk f(y, means, out2) (the forward accumulator),
//   g_means = scale dF/dmeans, g_out2 = scale dF/dout2  [N,K,S,Dobs],  g_w = scale dF/dw  [N,K].
// One warp per (n,k): it owns all S rows of the pair, so g_w needs no atomics; pure HBM stream.
template <typename T>
__global__ void __launch_bounds__(256)
decoder_loglike_bwd_kernel(int64_t N, int K, int S, int Dobs, int mode, const T* __restrict__ y,
                           const T* __restrict__ means, const T* __restrict__ out2, const T* __restrict__ w, T scale,
                           T* __restrict__ g_means, T* __restrict__ g_out2, T* __restrict__ g_w) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t nk = warp; nk < N * K; nk += nwarps) {
        const int64_t n = nk / K;
        const T wk = w[nk] * scale;
        const T* yr = y + n * Dobs;
        T fsum = T(0);
        for (int s = 0; s < S; ++s) {
            const int64_t row = nk * S + s;
            const T* o2 = out2 + row * Dobs;
            if (mode == 0) {
                const T* mu = means + row * Dobs;
                for (int d = lane; d < Dobs; d += 32) {
                    const T e = yr[d] - mu[d], v = o2[d], iv = T(1) / v;
                    fsum += e * e * iv + t_log(v + T(1e-8));
                    g_means[row * Dobs + d] = T(-2) * wk * e * iv;
                    g_out2[row * Dobs + d] = wk * (T(1) / (v + T(1e-8)) - e * e * iv * iv);
                }
            } else {
                for (int d = lane; d < Dobs; d += 32) {
                    const T t = -o2[d] * yr[d];
                    fsum -= t_softplus(t);
                    g_out2[row * Dobs + d] = wk * yr[d] / (T(1) + t_exp(-t));     // y * sigmoid(-logit y)
                }
            }
        }
        fsum = warp_sum(fsum);
        if (lane == 0 && g_w != nullptr) g_w[nk] = scale * fsum;
    }
}

template <typename T>
int decoder_loglike_bwd(int64_t N, int K, int S, int Dobs, int mode, const T* y, const T* means, const T* out2,
                        const T* w, double scale, T* g_means, T* g_out2, T* g_w, void* stream) {
    if (N < 0 || K <= 0 || S <= 0 || Dobs <= 0 || !y || !out2 || !w || !g_out2) return VMP_E_BADARG;
    if (mode != 0 && mode != 1) return VMP_E_BADMODE;
    if (mode == 0 && (!means || !g_means)) return VMP_E_BADARG;
    if (N == 0) return VMP_OK;
    int64_t grid = (N * K + 7) / 8;
    if (grid > 148 * 16) grid = 148 * 16;
    decoder_loglike_bwd_kernel<T><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(N, K, S, Dobs, mode, y, means, out2,
                                                                                    w, (T)scale, g_means, g_out2, g_w);
    return launch_status();
}

// ---- test-time metrics over the decoder outputs (losses.py) ----------------------------------------------------------
// One warp per (n,k) streams its S rows once and produces both reductions the reference's metrics are built from:
//   sq[n,k]  = 1/S sum_s sum_d m_d (t_nd - means_nksd)^2                       (weighted_mse, imputation_mse)
//   lse[n,k] = log sum_s exp( lw_nks + sum_d m_d log p(y_nd | means, out2) )   (diagonal_gaussian_logprob, bernoulli_logprob)
// m = missing-data mask (NULL = all ones), t = the MSE target (NULL = y), lw = per-sample log weights (NULL = 0).
template <typename T>
__global__ void __launch_bounds__(256)
decoder_metrics_kernel(int64_t N, int K, int S, int Dobs, int mode, const T* __restrict__ y, const T* __restrict__ tgt,
                       const T* __restrict__ means, const T* __restrict__ out2, const uint8_t* __restrict__ mask,
                       const T* __restrict__ lw, T* __restrict__ sq, T* __restrict__ lse) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t nk = warp; nk < N * K; nk += nwarps) {
        const int64_t n = nk / K;
        const T* yr = y + n * Dobs;
        const T* tr = (tgt != nullptr ? tgt : y) + n * Dobs;
        const uint8_t* mr = mask != nullptr ? mask + n * Dobs : nullptr;
        T sqs = T(0), mx = -CUDART_INF_F, acc = T(0);
        for (int s = 0; s < S; ++s) {
            const int64_t row = nk * S + s;
            const T* mu = means + row * Dobs;
            const T* o2 = out2 + row * Dobs;
            T se = T(0), lp = T(0);
            for (int d = lane; d < Dobs; d += 32) {
                if (mr != nullptr && !mr[d]) continue;
                const T e = tr[d] - mu[d];
                se = fma(e, e, se);
                if (mode == 0) {
                    const T ey = yr[d] - mu[d], v = o2[d];
                    lp -= T(0.5) * (ey * ey / v + t_log(v) + T(VMP_LOG_2PI));
                } else {
                    lp -= t_softplus(-o2[d] * yr[d]);
                }
            }
            se = warp_sum(se);
            lp = warp_sum(lp) + (lw != nullptr ? lw[row] : T(0));
            sqs += se;
            if (lp > mx) { acc = acc * t_exp(mx - lp) + T(1); mx = lp; }        // online log-sum-exp over samples
            else acc += t_exp(lp - mx);
        }
        if (lane == 0) {
            if (sq != nullptr) sq[nk] = sqs / T(S);
            if (lse != nullptr) lse[nk] = mx + t_log(acc);
        }
    }
}

template <typename T>
int decoder_metrics(int64_t N, int K, int S, int Dobs, int mode, const T* y, const T* tgt, const T* means, const T* out2,
                    const uint8_t* mask, const T* lw, T* sq, T* lse, void* stream) {
    if (N < 0 || K <= 0 || S <= 0 || Dobs <= 0 || !y || !means || !out2) return VMP_E_BADARG;
    if (mode != 0 && mode != 1) return VMP_E_BADMODE;
    if (N == 0) return VMP_OK;
    int64_t grid = (N * K + 7) / 8;
    if (grid > 148 * 16) grid = 148 * 16;
    decoder_metrics_kernel<T><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(N, K, S, Dobs, mode, y, tgt, means, out2,
                                                                                mask, lw, sq, lse);
    return launch_status();
}

// ---- general dense-natural-parameter Gaussian log density ---------------------------------------------------------
// One thread per (n,k): P = -2 eta2 = L L^T (packed, local memory), mu = P^-1 eta1,
//   log N(x) = -1/2 |L^T (x - mu)|^2 + sum log L_ii - D/2 log 2pi
// which equals x.eta1 + x^T eta2 x - D/2 log 2pi + 1/4 eta1^T eta2^-1 eta1 + 1/2 logdet(-2 eta2) (gaussian.py:52-61).
template <typename T>
__global__ void __launch_bounds__(128)
gaussian_logprob_kernel(int64_t N, int K, int S, int D, const T* __restrict__ x, const T* __restrict__ eta1,
                        const T* __restrict__ eta2, const T* __restrict__ log_w, T* __restrict__ out) {
    const int64_t pair = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= N * K) return;
    const int64_t n = pair / K;
    const int k = (int)(pair - n * K);
    T L[VMP_MAX_D * (VMP_MAX_D + 1) / 2], mu[VMP_MAX_D];
    const T* e2 = eta2 + pair * (int64_t)D * D;
    const T* e1 = eta1 + pair * (int64_t)D;
    T hld = T(0);
    for (int i = 0; i < D; ++i) {
        T sm = e1[i];
        for (int j = 0; j <= i; ++j) {
            T s = T(-2) * e2[i * D + j];
            for (int c = 0; c < j; ++c) s = fma(-L[i * (i + 1) / 2 + c], L[j * (j + 1) / 2 + c], s);
            if (j == i) {
                const T l = t_sqrt(s);
                L[i * (i + 1) / 2 + i] = l;
                hld += t_log(l);
                mu[i] = sm / l;
            } else {
                const T l = s / L[j * (j + 1) / 2 + j];
                L[i * (i + 1) / 2 + j] = l;
                sm = fma(-l, mu[j], sm);
            }
        }
    }
    for (int i = D - 1; i >= 0; --i) {                       // mu = L^-T (L^-1 eta1)
        T s = mu[i];
        for (int c = i + 1; c < D; ++c) s = fma(-L[c * (c + 1) / 2 + i], mu[c], s);
        mu[i] = s / L[i * (i + 1) / 2 + i];
    }
    const T cst = hld - T(0.5 * VMP_LOG_2PI) * T(D);
    const int ns = S > 0 ? S : 1;
    for (int s = 0; s < ns; ++s) {
        const T* xs = S > 0 ? x + (pair * S + s) * (int64_t)D : x + n * (int64_t)D;
        T q = T(0);
        for (int i = 0; i < D; ++i) {
            T t = T(0);
            for (int c = i; c < D; ++c) t = fma(L[c * (c + 1) / 2 + i], xs[c] - mu[c], t);
            q = fma(t, t, q);
        }
        T lp = cst - T(0.5) * q;
        if (S > 0) out[pair * S + s] = lp;
        else out[pair] = lp + (log_w != nullptr ? log_w[k] : T(0));
    }
}

template <typename T>
__global__ void normalise_rows_kernel(int64_t N, int K, T* __restrict__ lp) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    T* r = lp + n * K;
    T mx = r[0];
    for (int k = 1; k < K; ++k) mx = max(mx, r[k]);
    double se = 0.0;
    for (int k = 0; k < K; ++k) se += exp((double)(r[k] - mx));
    const T lse = mx + (T)log(se);
    for (int k = 0; k < K; ++k) r[k] -= lse;
}

template <typename T>
int gaussian_logprob_nat(int64_t N, int K, int S, int D, const T* x, const T* eta1, const T* eta2, const T* log_w,
                         T* out, void* stream) {
    if (N < 0 || K <= 0 || S < 0 || !x || !eta1 || !eta2 || !out) return VMP_E_BADARG;
    if (D < 1 || D > VMP_MAX_D) return VMP_E_BADDIM;
    if (N == 0) return VMP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t pairs = N * K;
    gaussian_logprob_kernel<T><<<(unsigned)((pairs + 127) / 128), 128, 0, st>>>(N, K, S, D, x, eta1, eta2, log_w, out);
    if (int e = launch_status()) return e;
    if (S == 0) {
        normalise_rows_kernel<T><<<(unsigned)((N + 127) / 128), 128, 0, st>>>(N, K, out);
        if (int e = launch_status()) return e;
    }
    return VMP_OK;
}

}  // namespace vmp

extern "C" {
int vmp_decoder_loglike_f32(int64_t N, int K, int S, int Dobs, int mode, const float* y, const float* means,
                            const float* out2, const float* w, double* acc, void* stream) {
    return vmp::decoder_loglike<float>(N, K, S, Dobs, mode, y, means, out2, w, acc, stream);
}
int vmp_decoder_loglike_f64(int64_t N, int K, int S, int Dobs, int mode, const double* y, const double* means,
                            const double* out2, const double* w, double* acc, void* stream) {
    return vmp::decoder_loglike<double>(N, K, S, Dobs, mode, y, means, out2, w, acc, stream);
}
int vmp_decoder_loglike_bwd_f32(int64_t N, int K, int S, int Dobs, int mode, const float* y, const float* means,
                                const float* out2, const float* w, double scale, float* g_means, float* g_out2,
                                float* g_w, void* stream) {
    return vmp::decoder_loglike_bwd<float>(N, K, S, Dobs, mode, y, means, out2, w, scale, g_means, g_out2, g_w, stream);
}
int vmp_decoder_loglike_bwd_f64(int64_t N, int K, int S, int Dobs, int mode, const double* y, const double* means,
                                const double* out2, const double* w, double scale, double* g_means, double* g_out2,
                                double* g_w, void* stream) {
    return vmp::decoder_loglike_bwd<double>(N, K, S, Dobs, mode, y, means, out2, w, scale, g_means, g_out2, g_w, stream);
}
int vmp_decoder_metrics_f32(int64_t N, int K, int S, int Dobs, int mode, const float* y, const float* target,
                            const float* means, const float* out2, const uint8_t* mask, const float* log_w_nks,
                            float* sq, float* lse, void* stream) {
    return vmp::decoder_metrics<float>(N, K, S, Dobs, mode, y, target, means, out2, mask, log_w_nks, sq, lse, stream);
}
int vmp_decoder_metrics_f64(int64_t N, int K, int S, int Dobs, int mode, const double* y, const double* target,
                            const double* means, const double* out2, const uint8_t* mask, const double* log_w_nks,
                            double* sq, double* lse, void* stream) {
    return vmp::decoder_metrics<double>(N, K, S, Dobs, mode, y, target, means, out2, mask, log_w_nks, sq, lse, stream);
}
int vmp_gaussian_logprob_nat_f32(int64_t N, int K, int S, int D, const float* x, const float* eta1, const float* eta2,
                                 const float* log_w, float* out, void* stream) {
    return vmp::gaussian_logprob_nat<float>(N, K, S, D, x, eta1, eta2, log_w, out, stream);
}
int vmp_gaussian_logprob_nat_f64(int64_t N, int K, int S, int D, const double* x, const double* eta1,
                                 const double* eta2, const double* log_w, double* out, void* stream) {
    return vmp::gaussian_logprob_nat<double>(N, K, S, D, x, eta1, eta2, log_w, out, stream);
}
}
