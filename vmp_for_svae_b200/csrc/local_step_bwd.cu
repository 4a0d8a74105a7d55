// local_step_bwd.cu — reverse pass of the fused local step: gradients of
//     sum(gx * x_k_samples) + sum(glr * log_r) + greg * regulariser
// w.r.t. the encoder potentials (eta1, eta2_diag) and the raw recognition-GMM parameters (eta1_phi2, L_raw, pi_raw);
// theta is behind the reference's tf.stop_gradient (svae.py:211-214).  This is what opt.compute_gradients(-elbo)
// (experiments.py:232) back-propagates through svae.e_step / compute_elbo.  The per-pair formulas are those of
// oracle/backward.py::backward_closed_form (checked against torch.autograd in tests/):
//   Sig = P~^-1, mu~ = Sig eta~, u_s = L^-T eps_s, x_s = mu~ + u_s
//   Gx_s = gx_s + greg/S * r W^T W (x_s - m);  glr' = glr + greg r (T + 1);  s_bar = glr' - r sum_k glr'
//   P~_bar = -sym(v mu~^T) + sym(L^-T Phi(L^T L_bar) L^-1) + 1/2 (greg r - s_bar) Sig - 1/2 s_bar b b^T
//            v = Sig sum_s Gx_s, L_bar = -tril(sum_s u_s (L^-1 Gx_s)^T), b = Sig P2 d
//   d_bar = -s_bar P2 (d - b);  P2_bar = P~_bar - 1/2 s_bar (d d^T - d b^T - b d^T)  (+ K-sized terms in the epilogue)
// One thread per (point, component) pair, latent dimension <= 16 (the reference's training shapes are D = 2 and 6);
// per-point sums go through shared memory, per-component sums through double atomics, and a K-sized epilogue
// kernel maps (P2_bar, mu2_bar, s_sum) back to (eta1_phi2, L_raw, pi_raw).
#include "common.cuh"

namespace vmp {

constexpr int BWD_MAX_D = 16;
// per-component accumulator: P2_bar D^2 | h2_bar D | mu2_bar D | s_sum | W_bar D^2 | m_bar D | cden_bar
__host__ __device__ inline int bwd_klen(int D) { return 2 * D * D + 3 * D + 2; }

template <typename T, int DT, bool TH>
__global__ void __launch_bounds__(256) local_step_bwd_kernel(int64_t N, int K, int Drt, int S, int PTS, int den_mode, const T* __restrict__ eta1,
                      const T* __restrict__ eta2d, const T* __restrict__ phi_rec, const T* __restrict__ theta_rec,
                      const T* __restrict__ noise, uint64_t seed, const T* __restrict__ log_r,
                      const T* __restrict__ gx, const T* __restrict__ glr, T greg_host, const T* __restrict__ greg_dev,
                      T* __restrict__ eta1_bar, T* __restrict__ eta2d_bar, double* __restrict__ kacc) {
    constexpr int DM = DT ? DT : BWD_MAX_D;
    const int D = DT ? DT : Drt;
    const T greg = greg_dev != nullptr ? *greg_dev : greg_host;
    extern __shared__ __align__(8) unsigned char smraw[];
    double* pacc = reinterpret_cast<double*>(smraw);                 // [PTS][2*D]  eta1_bar | p1_bar per point
    T* gsum = reinterpret_cast<T*>(pacc + (size_t)PTS * 2 * D);      // [PTS]       sum_k glr'
    const int tid = threadIdx.x;
    const int64_t pt0 = (int64_t)blockIdx.x * PTS;
    const int npts = (int)min((int64_t)PTS, N - pt0);
    const int pl = tid / K, k = tid - pl * K;
    const bool act = pl < npts;
    const int64_t n = pt0 + (act ? pl : 0);
    for (int e = tid; e < PTS * 2 * D; e += blockDim.x) pacc[e] = 0.0;
    for (int e = tid; e < PTS; e += blockDim.x) gsum[e] = T(0);
    __syncthreads();

    const int plen = phi_record_len(D), tlen = theta_record_len(D), kl = bwd_klen(D);
    const T* prec = phi_rec + (size_t)k * plen;
    const T* P2 = prec;
    const T* trec = theta_rec + (size_t)k * tlen;
    const T* W = trec;
    const T* mth = trec + D * D;

    T Lc[DM * DM], Li[DM * DM], Sg[DM * DM];      // L (lower), L^-1 (lower), Sig = P~^-1 (full)
    T p1[DM], mu1[DM], dv[DM], bv[DM], mut[DM], gmu[DM], Lb[DM * DM];
    T r = T(0), glrp = T(0), hldv = T(0);
    T Wb[TH ? DM * DM : 1], mb[TH ? DM : 1];      // d/dW (lower), d/dm of the regulariser (Student-t theta is trained)
    if (act) {
#pragma unroll
        for (int i = 0; i < D; ++i) {
            p1[i] = T(-2) * eta2d[n * D + i];
            mu1[i] = eta1[n * D + i] / p1[i];
            dv[i] = mu1[i] - prec[D * D + i];
        }
        // L = chol(P2 + diag p1)
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                T s = P2[i * D + j] + (i == j ? p1[i] : T(0));
#pragma unroll
                for (int c = 0; c < j; ++c) s = fma(-Lc[i * D + c], Lc[j * D + c], s);
                Lc[i * D + j] = (i == j) ? t_sqrt(s) : s / Lc[j * D + j];
            }
        // L^-1 (lower), column by column
#pragma unroll
        for (int j = 0; j < D; ++j) {
            Li[j * D + j] = T(1) / Lc[j * D + j];
#pragma unroll
            for (int i = j + 1; i < D; ++i) {
                T s = T(0);
#pragma unroll
                for (int c = j; c < i; ++c) s = fma(Lc[i * D + c], Li[c * D + j], s);
                Li[i * D + j] = -s / Lc[i * D + i];
            }
        }
        // Sig = L^-T L^-1
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                T s = T(0);
#pragma unroll
                for (int c = i; c < D; ++c) s = fma(Li[c * D + i], Li[c * D + j], s);
                Sg[i * D + j] = s;
                Sg[j * D + i] = s;
            }
        hldv = T(0);
#pragma unroll
        for (int i = 0; i < D; ++i) hldv += t_log(Lc[i * D + i]);
        // mu~ = Sig (eta1 + h2) ; b = Sig P2 d
        T gq[DM];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            T s = T(0);
#pragma unroll
            for (int c = 0; c < D; ++c) s = fma(P2[i * D + c], dv[c], s);
            gq[i] = s;
        }
#pragma unroll
        for (int i = 0; i < D; ++i) {
            T s0 = T(0), s1 = T(0);
#pragma unroll
            for (int c = 0; c < D; ++c) {
                s0 = fma(Sg[i * D + c], eta1[n * D + c] + prec[D * D + D + c], s0);
                s1 = fma(Sg[i * D + c], gq[c], s1);
            }
            mut[i] = s0;
            bv[i] = s1;
            gmu[i] = T(0);
        }
#pragma unroll
        for (int e = 0; e < D * D; ++e) Lb[e] = T(0);
        if (TH) {
#pragma unroll
            for (int e = 0; e < D * D; ++e) Wb[e] = T(0);
#pragma unroll
            for (int i = 0; i < D; ++i) mb[i] = T(0);
        }
        r = t_exp(log_r[n * K + k]);
        // samples: u_s, x_s, Gx_s, t_s ; accumulate gmu, L_bar, T
        T Tsum = T(0);
        for (int s = 0; s < S; ++s) {
            T eps[DM], u[DM], Gx[DM], tv[DM], wx[DM];
            const uint64_t pair = (uint64_t)n * K + k;
            T e2 = T(0);
#pragma unroll
            for (int i = 0; i < D; ++i) {
                eps[i] = noise != nullptr ? noise[(pair * D + i) * (uint64_t)S + s]
                                          : (T)philox_normal1(seed, pair, (uint32_t)s, (uint32_t)i);
                e2 = fma(eps[i], eps[i], e2);
            }
#pragma unroll
            for (int ii = 0; ii < D; ++ii) {                       // u = L^-T eps
                const int i = D - 1 - ii;
                T sacc = eps[i];
#pragma unroll
                for (int c = i + 1; c < D; ++c) sacc = fma(-Lc[c * D + i], u[c], sacc);
                u[i] = sacc / Lc[i * D + i];
            }
            T q2 = T(0);
#pragma unroll
            for (int i = 0; i < D; ++i) {                          // wx = W (x - m)
                T sacc = T(0);
#pragma unroll
                for (int c = 0; c <= i; ++c) sacc = fma(W[i * D + c], mut[c] + u[c] - mth[c], sacc);
                wx[i] = sacc;
                q2 = fma(sacc, sacc, q2);
            }
            const T nu = trec[D * D + D + 1];
            const T coef = den_mode == VMP_DEN_GAUSS ? T(1) : (nu + T(D)) / (nu + q2);      // d(-den)/d(q2/2)
            const T* gxs = gx + ((pair * S) + s) * (uint64_t)D;
#pragma unroll
            for (int i = 0; i < D; ++i) {                          // Gx = gx + greg/S r W^T wx
                T sacc = T(0);
#pragma unroll
                for (int c = i; c < D; ++c) sacc = fma(W[c * D + i], wx[c], sacc);
                const T gden = greg / T(S) * r * coef * sacc;        // -greg r/S d den/d x
                Gx[i] = gxs[i] + gden;
                gmu[i] += Gx[i];
                if (TH) mb[i] -= gden;
            }
            if (TH) {
                const T cf = greg / T(S) * r * coef;
#pragma unroll
                for (int i = 0; i < D; ++i)
#pragma unroll
                    for (int c = 0; c <= i; ++c) Wb[i * D + c] = fma(cf * wx[i], mut[c] + u[c] - mth[c], Wb[i * D + c]);
            }
#pragma unroll
            for (int i = 0; i < D; ++i) {                          // t = L^-1 Gx
                T sacc = T(0);
#pragma unroll
                for (int c = 0; c <= i; ++c) sacc = fma(Li[i * D + c], Gx[c], sacc);
                tv[i] = sacc;
            }
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) Lb[i * D + j] = fma(-u[i], tv[j], Lb[i * D + j]);
            const T num = T(-0.5) * e2 + hldv - T(0.5 * VMP_LOG_2PI) * T(D) + log_r[n * K + k];
            const T den = den_mode == VMP_DEN_GAUSS ? trec[D * D + D] - T(0.5) * q2
                                                    : trec[D * D + D] - T(0.5) * (nu + T(D)) * t_log1p(q2 / nu);
            Tsum += num - den;
        }
        glrp = glr[n * K + k] + greg * r * (Tsum / T(S) + T(1));
        atomicAdd(&gsum[pl], glrp);
    }
    __syncthreads();
    if (act) {
    const T s_bar = glrp - r * gsum[pl];
    const T hld_bar = greg * r - s_bar;

    // v = Sig gmu ; Pt_bar (full symmetric) in Pb
    T v[DM], Pb[DM * DM];
#pragma unroll
    for (int i = 0; i < D; ++i) {
        T sacc = T(0);
#pragma unroll
        for (int c = 0; c < D; ++c) sacc = fma(Sg[i * D + c], gmu[c], sacc);
        v[i] = sacc;
    }
    // Phi = tril(L^T L_bar), diagonal halved
    T Ph[DM * DM];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            T sacc = T(0);
#pragma unroll
            for (int c = i; c < D; ++c) sacc = fma(Lc[c * D + i], Lb[c * D + j], sacc);
            Ph[i * D + j] = (i == j) ? T(0.5) * sacc : sacc;
        }
    // Sm = L^-T Phi L^-1 ;  X = Phi Li (lower x lower -> lower), Sm = Li^T X
    T X[DM * DM];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            T sacc = T(0);
#pragma unroll
            for (int c = j; c <= i; ++c) sacc = fma(Ph[i * D + c], Li[c * D + j], sacc);
            X[i * D + j] = sacc;
        }
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
            T sacc = T(0);
#pragma unroll
            for (int c = (i > j ? i : j); c < D; ++c) sacc = fma(Li[c * D + i], X[c * D + j], sacc);   // X[c][j]: j <= c
            Pb[i * D + j] = sacc;                                   // Sm[i][j]
        }
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            const T sm = T(0.5) * (Pb[i * D + j] + Pb[j * D + i]);
            const T val = sm - T(0.5) * (v[i] * mut[j] + mut[i] * v[j]) + T(0.5) * hld_bar * Sg[i * D + j]
                          - T(0.5) * s_bar * bv[i] * bv[j];
            Pb[i * D + j] = val;
            Pb[j * D + i] = val;
        }
    // d_bar = -s_bar P2 (d - b)
    T dbar[DM];
#pragma unroll
    for (int i = 0; i < D; ++i) {
        T sacc = T(0);
#pragma unroll
        for (int c = 0; c < D; ++c) sacc = fma(P2[i * D + c], dv[c] - bv[c], sacc);
        dbar[i] = -s_bar * sacc;
    }
    // per-point sums: eta1_bar += v + d_bar / p1 ; p1_bar += diag(Pt_bar) - d_bar * eta1 / p1^2
#pragma unroll
    for (int i = 0; i < D; ++i) {
        atomicAdd(&pacc[(size_t)pl * 2 * D + i], (double)(v[i] + dbar[i] / p1[i]));
        atomicAdd(&pacc[(size_t)pl * 2 * D + D + i], (double)(Pb[i * D + i] - dbar[i] * mu1[i] / p1[i]));
    }
    // per-component sums (double atomics): P2_bar | h2_bar (v) | mu2_bar (-d_bar) | s_sum
    double* ka = kacc + (size_t)k * kl;
#pragma unroll
    for (int i = 0; i < D; ++i) {
#pragma unroll
        for (int j = 0; j < D; ++j) {
            const T val = Pb[i * D + j] - T(0.5) * s_bar * (dv[i] * dv[j] - dv[i] * bv[j] - bv[i] * dv[j]);
            atomicAdd(ka + i * D + j, (double)val);
        }
        atomicAdd(ka + D * D + i, (double)v[i]);
        atomicAdd(ka + D * D + D + i, (double)(-dbar[i]));
    }
    atomicAdd(ka + D * D + 2 * D, (double)s_bar);
    if (TH) {
        double* kt = ka + D * D + 2 * D + 1;
#pragma unroll
        for (int i = 0; i < D; ++i) {
#pragma unroll
            for (int c = 0; c <= i; ++c) atomicAdd(kt + i * D + c, (double)Wb[i * D + c]);
            atomicAdd(kt + D * D + i, (double)mb[i]);
        }
        atomicAdd(kt + D * D + D, (double)(-greg * r));
    }
    }
    __syncthreads();
    for (int e = tid; e < npts * D; e += blockDim.x) {             // per-point results out of the shared accumulators
        const int q = e / D, i = e - q * D;
        eta1_bar[(pt0 + q) * D + i] = (T)pacc[(size_t)q * 2 * D + i];
        eta2d_bar[(pt0 + q) * D + i] = (T)(-2.0 * pacc[(size_t)q * 2 * D + D + i]);
    }
}

// ---- D > 16: one CTA per POINT, the D x D matrices of the current (point, component) pair live in shared memory and every
// step (Cholesky, triangular inverse, the Murray reverse products) is block-cooperative.  Two passes over the components: the
// first only evaluates glr' = glr + greg r (mean_s(num - den) + 1) for the log-soft-max reverse (needs sum_k glr'), the second
// recomputes the factorisation and finishes as the thread-per-pair kernel above does.  Same formulas, same accumulators.
constexpr int BWB_THREADS = 256;

template <typename T>
__device__ void bwb_chol(T* A, int D, int ld) {                      // in place, lower; all threads
    for (int j = 0; j < D; ++j) {
        __syncthreads();
        const T djj = t_sqrt(A[j * ld + j]);
        __syncthreads();
        if (threadIdx.x == 0) A[j * ld + j] = djj;
        const T inv = T(1) / djj;
        for (int i = j + 1 + threadIdx.x; i < D; i += blockDim.x) A[i * ld + j] *= inv;
        __syncthreads();
        const int m = D - j - 1;
        for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
            const int i = j + 1 + e / m, c = j + 1 + e % m;
            if (c <= i) A[i * ld + c] = fma(-A[i * ld + j], A[c * ld + j], A[i * ld + c]);
        }
    }
    __syncthreads();
}

template <typename T, bool TH>
__global__ void __launch_bounds__(BWB_THREADS)
local_step_bwd_block_kernel(int64_t N, int K, int D, int S, int den_mode, const T* __restrict__ eta1,
                            const T* __restrict__ eta2d, const T* __restrict__ phi_rec, const T* __restrict__ theta_rec,
                            const T* __restrict__ noise, uint64_t seed, const T* __restrict__ log_r, const T* __restrict__ gx,
                            const T* __restrict__ glr, T greg_host, const T* __restrict__ greg_dev, T* __restrict__ eta1_bar,
                            T* __restrict__ eta2d_bar, double* __restrict__ kacc) {
    const T greg = greg_dev != nullptr ? *greg_dev : greg_host;
    const int ld = D + 1, tid = threadIdx.x, nth = blockDim.x;
    const int64_t n = blockIdx.x;
    extern __shared__ __align__(16) unsigned char smraw[];
    double* pacc = reinterpret_cast<double*>(smraw);                 // [2D]
    T* Lc = reinterpret_cast<T*>(pacc + 2 * D);                      // chol(P~)
    T* Li = Lc + D * ld;                                             // L^-1
    T* Sg = Li + D * ld;                                             // P~^-1
    T* Lb = Sg + D * ld;                                             // L_bar -> X
    T* B5 = Lb + D * ld;                                             // Phi -> Sm -> Pt_bar
    T* Wb = B5 + D * ld;                                             // d/dW (TH only; else zero-sized use)
    T* vecs = Wb + (TH ? D * ld : 0);
    T *p1 = vecs, *mu1 = p1 + D, *dv = mu1 + D, *gq = dv + D, *mut = gq + D, *bv = mut + D, *gmu = bv + D, *eps = gmu + D,
      *u = eps + D, *xm = u + D, *wx = xm + D, *Gx = wx + D, *tv = Gx + D, *v = tv + D, *dbar = v + D, *mb = dbar + D,
      *hh = mb + D, *glrp = hh + D;                                  // glrp[K]
    __shared__ double red[32];
    __shared__ T sh_scalar[4];
    const int plen = phi_record_len(D), tlen = theta_record_len(D), kl = bwd_klen(D);
    for (int i = tid; i < D; i += nth) {
        p1[i] = T(-2) * eta2d[n * D + i];
        mu1[i] = eta1[n * D + i] / p1[i];
    }
    for (int i = tid; i < 2 * D; i += nth) pacc[i] = 0.0;
    __syncthreads();
    T gsum = T(0);
    for (int pass = 0; pass < 2; ++pass) {
        for (int k = 0; k < K; ++k) {
            const T* prec = phi_rec + (size_t)k * plen;
            const T* P2 = prec;
            const T* trec = theta_rec + (size_t)k * tlen;
            const T* W = trec;
            const T* mth = trec + D * D;
            __syncthreads();
            for (int e = tid; e < D * D; e += nth) {
                const int i = e / D, j = e - i * D;
                Lc[i * ld + j] = P2[e] + (i == j ? p1[i] : T(0));
            }
            for (int i = tid; i < D; i += nth) {
                dv[i] = mu1[i] - prec[D * D + i];
                hh[i] = eta1[n * D + i] + prec[D * D + D + i];
                gmu[i] = T(0);
                if (TH) mb[i] = T(0);
            }
            bwb_chol(Lc, D, ld);
            for (int j = tid; j < D; j += nth) {                      // Li = L^-1, thread per column
                for (int i = 0; i < j; ++i) Li[i * ld + j] = T(0);
                Li[j * ld + j] = T(1) / Lc[j * ld + j];
                for (int i = j + 1; i < D; ++i) {
                    T sacc = T(0);
                    for (int c = j; c < i; ++c) sacc = fma(Lc[i * ld + c], Li[c * ld + j], sacc);
                    Li[i * ld + j] = -sacc / Lc[i * ld + i];
                }
            }
            double hl = 0.0;
            for (int i = tid; i < D; i += nth) hl += (double)t_log(Lc[i * ld + i]);
            hl = block_sum(hl, red);
            if (tid == 0) sh_scalar[0] = (T)hl;
            __syncthreads();
            const T hldv = sh_scalar[0];
            // Sig = L^-T L^-1 (pass 1 needs the matrix; pass 0 only mu~ = Sig (eta1 + h2), done with it as well)
            for (int e = tid; e < D * D; e += nth) {
                const int i = e / D, j = e - i * D;
                if (j > i) continue;
                T sacc = T(0);
                for (int c = i; c < D; ++c) sacc = fma(Li[c * ld + i], Li[c * ld + j], sacc);
                Sg[i * ld + j] = sacc;
                Sg[j * ld + i] = sacc;
            }
            for (int i = tid; i < D; i += nth) {
                T sacc = T(0);
                for (int c = 0; c < D; ++c) sacc = fma(P2[i * D + c], dv[c], sacc);
                gq[i] = sacc;
            }
            __syncthreads();
            for (int i = tid; i < D; i += nth) {
                T s0 = T(0), s1 = T(0);
                for (int c = 0; c < D; ++c) {
                    s0 = fma(Sg[i * ld + c], hh[c], s0);
                    s1 = fma(Sg[i * ld + c], gq[c], s1);
                }
                mut[i] = s0;
                bv[i] = s1;
            }
            if (pass == 1) {
                for (int e = tid; e < D * ld; e += nth) { Lb[e] = T(0); if (TH) Wb[e] = T(0); }
            }
            __syncthreads();
            const T r = t_exp(log_r[n * K + k]);
            const T nu = trec[D * D + D + 1];
            const uint64_t pair = (uint64_t)n * K + k;
            T Tsum = T(0);
            for (int s = 0; s < S; ++s) {
                for (int i = tid; i < D; i += nth)
                    eps[i] = noise != nullptr ? noise[(pair * D + i) * (uint64_t)S + s]
                                              : (T)philox_normal1(seed, pair, (uint32_t)s, (uint32_t)i);
                __syncthreads();
                for (int i = tid; i < D; i += nth) {                  // u = L^-T eps = Li^T eps ; xm = mu~ + u - m_theta
                    T sacc = T(0);
                    for (int c = i; c < D; ++c) sacc = fma(Li[c * ld + i], eps[c], sacc);
                    u[i] = sacc;
                    xm[i] = mut[i] + sacc - mth[i];
                }
                __syncthreads();
                double q2d = 0.0, e2d = 0.0;
                for (int i = tid; i < D; i += nth) {                  // wx = W (x - m)
                    T sacc = T(0);
                    for (int c = 0; c <= i; ++c) sacc = fma(W[i * D + c], xm[c], sacc);
                    wx[i] = sacc;
                    q2d += (double)sacc * (double)sacc;
                    e2d += (double)eps[i] * (double)eps[i];
                }
                q2d = block_sum(q2d, red);
                e2d = block_sum(e2d, red);
                if (tid == 0) { sh_scalar[1] = (T)q2d; sh_scalar[2] = (T)e2d; }
                __syncthreads();
                const T q2 = sh_scalar[1], e2 = sh_scalar[2];
                const T num = T(-0.5) * e2 + hldv - T(0.5 * VMP_LOG_2PI) * T(D) + log_r[n * K + k];
                const T den = den_mode == VMP_DEN_GAUSS ? trec[D * D + D] - T(0.5) * q2
                                                        : trec[D * D + D] - T(0.5) * (nu + T(D)) * t_log1p(q2 / nu);
                Tsum += num - den;
                if (pass == 1) {
                    const T coef = den_mode == VMP_DEN_GAUSS ? T(1) : (nu + T(D)) / (nu + q2);
                    const T cf = greg / T(S) * r * coef;
                    const T* gxs = gx + ((pair * S) + s) * (uint64_t)D;
                    for (int i = tid; i < D; i += nth) {              // Gx = gx + greg/S r coef W^T wx
                        T sacc = T(0);
                        for (int c = i; c < D; ++c) sacc = fma(W[c * D + i], wx[c], sacc);
                        const T gden = cf * sacc;
                        Gx[i] = gxs[i] + gden;
                        gmu[i] += Gx[i];
                        if (TH) mb[i] -= gden;
                    }
                    __syncthreads();
                    for (int i = tid; i < D; i += nth) {              // t = L^-1 Gx
                        T sacc = T(0);
                        for (int c = 0; c <= i; ++c) sacc = fma(Li[i * ld + c], Gx[c], sacc);
                        tv[i] = sacc;
                    }
                    __syncthreads();
                    for (int e = tid; e < D * D; e += nth) {
                        const int i = e / D, j = e - i * D;
                        if (j > i) continue;
                        Lb[i * ld + j] = fma(-u[i], tv[j], Lb[i * ld + j]);
                        if (TH) Wb[i * ld + j] = fma(cf * wx[i], xm[j], Wb[i * ld + j]);
                    }
                }
                __syncthreads();
            }
            if (pass == 0) {
                if (tid == 0) glrp[k] = glr[n * K + k] + greg * r * (Tsum / T(S) + T(1));
                continue;
            }
            // ---------------- finish (pass 1)
            const T s_bar = glrp[k] - r * gsum;
            const T hld_bar = greg * r - s_bar;
            for (int i = tid; i < D; i += nth) {                      // v = Sig gmu
                T sacc = T(0);
                for (int c = 0; c < D; ++c) sacc = fma(Sg[i * ld + c], gmu[c], sacc);
                v[i] = sacc;
            }
            for (int e = tid; e < D * D; e += nth) {                  // Phi = tril(L^T L_bar), diagonal halved -> B5
                const int i = e / D, j = e - i * D;
                if (j > i) { B5[i * ld + j] = T(0); continue; }
                T sacc = T(0);
                for (int c = i; c < D; ++c) sacc = fma(Lc[c * ld + i], Lb[c * ld + j], sacc);
                B5[i * ld + j] = (i == j) ? T(0.5) * sacc : sacc;
            }
            __syncthreads();
            for (int e = tid; e < D * D; e += nth) {                  // X = Phi L^-1 (lower) -> Lb
                const int i = e / D, j = e - i * D;
                T sacc = T(0);
                if (j <= i)
                    for (int c = j; c <= i; ++c) sacc = fma(B5[i * ld + c], Li[c * ld + j], sacc);
                Lb[i * ld + j] = sacc;
            }
            __syncthreads();
            for (int e = tid; e < D * D; e += nth) {                  // Sm = L^-T X -> B5
                const int i = e / D, j = e - i * D;
                T sacc = T(0);
                for (int c = (i > j ? i : j); c < D; ++c) sacc = fma(Li[c * ld + i], Lb[c * ld + j], sacc);
                B5[i * ld + j] = sacc;
            }
            __syncthreads();
            for (int e = tid; e < D * D; e += nth) {                  // Pt_bar (symmetric) in place
                const int i = e / D, j = e - i * D;
                if (j > i) continue;
                const T sm = T(0.5) * (B5[i * ld + j] + B5[j * ld + i]);
                const T val = sm - T(0.5) * (v[i] * mut[j] + mut[i] * v[j]) + T(0.5) * hld_bar * Sg[i * ld + j] -
                              T(0.5) * s_bar * bv[i] * bv[j];
                B5[i * ld + j] = val;
                B5[j * ld + i] = val;
            }
            for (int i = tid; i < D; i += nth) {                      // d_bar = -s_bar P2 (d - b)
                T sacc = T(0);
                for (int c = 0; c < D; ++c) sacc = fma(P2[i * D + c], dv[c] - bv[c], sacc);
                dbar[i] = -s_bar * sacc;
            }
            __syncthreads();
            double* ka = kacc + (size_t)k * kl;
            for (int i = tid; i < D; i += nth) {
                pacc[i] += (double)(v[i] + dbar[i] / p1[i]);
                pacc[D + i] += (double)(B5[i * ld + i] - dbar[i] * mu1[i] / p1[i]);
                atomicAdd(ka + D * D + i, (double)v[i]);
                atomicAdd(ka + D * D + D + i, (double)(-dbar[i]));
                if (TH) atomicAdd(ka + D * D + 2 * D + 1 + D * D + i, (double)mb[i]);
            }
            for (int e = tid; e < D * D; e += nth) {
                const int i = e / D, j = e - i * D;
                const T val = B5[i * ld + j] - T(0.5) * s_bar * (dv[i] * dv[j] - dv[i] * bv[j] - bv[i] * dv[j]);
                atomicAdd(ka + e, (double)val);
                if (TH && j <= i) atomicAdd(ka + D * D + 2 * D + 1 + e, (double)Wb[i * ld + j]);
            }
            if (tid == 0) {
                atomicAdd(ka + D * D + 2 * D, (double)s_bar);
                if (TH) atomicAdd(ka + D * D + 2 * D + 1 + D * D + D, (double)(-greg * r));
            }
        }
        if (pass == 0) {
            __syncthreads();
            double gs = 0.0;
            for (int k = tid; k < K; k += nth) gs += (double)glrp[k];
            gs = block_sum(gs, red);
            if (tid == 0) sh_scalar[3] = (T)gs;
            __syncthreads();
            gsum = sh_scalar[3];
        }
    }
    __syncthreads();
    for (int i = tid; i < D; i += nth) {
        eta1_bar[n * D + i] = (T)pacc[i];
        eta2d_bar[n * D + i] = (T)(-2.0 * pacc[D + i]);
    }
}

template <typename T>
static size_t bwb_smem(int K, int D, bool th) {
    return sizeof(double) * 2 * D + sizeof(T) * ((size_t)(th ? 6 : 5) * D * (D + 1) + 17 * (size_t)D + K) + 32;
}

// K-sized epilogue: (P2_bar, h2_bar, mu2_bar, s_sum) -> (eta1_phi2_bar, L_raw_bar, pi_raw_bar)
template <typename T>
__global__ void __launch_bounds__(128)
local_step_bwd_epilogue_kernel(int K, int D, const T* __restrict__ eta1_phi2, const T* __restrict__ L_raw,
                               const T* __restrict__ pi_raw, const double* __restrict__ kacc,
                               T* __restrict__ h2_bar, T* __restrict__ L_raw_bar, T* __restrict__ pi_raw_bar,
                               T* __restrict__ theta_rec_bar) {
    extern __shared__ double sm[];
    const int ld = D + 1;
    double* L2 = sm;                 // D x ld
    double* Wi = L2 + D * ld;        // L2^-1
    double* Pi = Wi + D * ld;        // P2^-1
    double* Pb = Pi + D * ld;        // P2_bar (full)
    double* vec = Pb + D * ld;       // mu2 | w2 | mu2_bar
    double* red = vec + 3 * D;
    const int k = blockIdx.x, kl = bwd_klen(D);
    const double* ka = kacc + (size_t)k * kl;
    for (int e = threadIdx.x; e < D * D; e += blockDim.x) {
        const int i = e / D, j = e % D;
        double v = 0.0;
        if (j < i) v = (double)L_raw[(size_t)k * D * D + e];
        else if (j == i) v = t_softplus<double>((double)L_raw[(size_t)k * D * D + e]);
        L2[i * ld + j] = v;
        Pb[i * ld + j] = ka[e];
    }
    __syncthreads();
    for (int j = threadIdx.x; j < D; j += blockDim.x) {             // Wi = L2^-1, column j
        for (int i = 0; i < j; ++i) Wi[i * ld + j] = 0.0;
        Wi[j * ld + j] = 1.0 / L2[j * ld + j];
        for (int i = j + 1; i < D; ++i) {
            double s = 0.0;
            for (int c = j; c < i; ++c) s += L2[i * ld + c] * Wi[c * ld + j];
            Wi[i * ld + j] = -s / L2[i * ld + i];
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < D * D; e += blockDim.x) {         // P2^-1 = Wi^T Wi
        const int i = e / D, j = e % D;
        const int m = i > j ? i : j;
        double s = 0.0;
        for (int c = m; c < D; ++c) s += Wi[c * ld + i] * Wi[c * ld + j];
        Pi[i * ld + j] = s;
    }
    __syncthreads();
    const double s_sum = ka[D * D + 2 * D];
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        double mu2 = 0.0, w2 = 0.0;
        for (int c = 0; c < D; ++c) {
            mu2 += Pi[i * ld + c] * (double)eta1_phi2[(size_t)k * D + c];
            w2 += Pi[i * ld + c] * ka[D * D + D + c];
        }
        vec[i] = mu2;
        vec[D + i] = w2;
        h2_bar[(size_t)k * D + i] = (T)(ka[D * D + i] + w2);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < D * D; e += blockDim.x) {
        const int i = e / D, j = e % D;
        Pb[i * ld + j] += 0.5 * s_sum * Pi[i * ld + j] - 0.5 * (vec[D + i] * vec[j] + vec[i] * vec[D + j]);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < D * D; e += blockDim.x) {         // L2_bar = tril((Pb + Pb^T) L2), softplus' on the diagonal
        const int i = e / D, j = e % D;
        double g = 0.0;
        if (j <= i) {
            for (int c = j; c < D; ++c) g += (Pb[i * ld + c] + Pb[c * ld + i]) * L2[c * ld + j];
            if (j == i) g *= 1.0 / (1.0 + exp(-(double)L_raw[(size_t)k * D * D + e]));
        }
        L_raw_bar[(size_t)k * D * D + e] = (T)g;
    }
    if (theta_rec_bar != nullptr) {                                   // W_bar | m_bar | cden_bar, 0, 0, 0
        const int tlen = theta_record_len(D);
        for (int e = threadIdx.x; e < tlen; e += blockDim.x)
            theta_rec_bar[(size_t)k * tlen + e] = e < D * D + D + 1 ? (T)ka[D * D + 2 * D + 1 + e] : T(0);
    }
    // pi_raw_bar = s_sum - softmax(pi_raw) * sum_k s_sum
    double mx = -CUDART_INF, tot = 0.0;
    for (int j = threadIdx.x; j < K; j += blockDim.x) mx = fmax(mx, (double)pi_raw[j]);
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmax(mx, red[w]);
    __syncthreads();
    double se = 0.0;
    for (int j = threadIdx.x; j < K; j += blockDim.x) {
        se += exp((double)pi_raw[j] - mx);
        tot += kacc[(size_t)j * kl + D * D + 2 * D];
    }
    se = block_sum(se, red);
    __shared__ double bse;
    if (threadIdx.x == 0) bse = se;
    __syncthreads();
    tot = block_sum(tot, red);
    if (threadIdx.x == 0) pi_raw_bar[k] = (T)(s_sum - exp((double)pi_raw[k] - mx) / bse * tot);
}

template <typename T, int DT, bool TH>
static cudaError_t launch_bwd_main(int64_t N, int K, int D, int S, int den_mode, const T* eta1, const T* eta2d,
                                   const T* phi_rec, const T* theta_rec, const T* noise, uint64_t seed, const T* log_r,
                                   const T* gx, const T* glr, T greg, const T* greg_dev, T* eta1_bar, T* eta2d_bar,
                                   double* kacc, cudaStream_t st) {
    const int PTS = K >= 128 ? 1 : 128 / K;
    const int threads = ((PTS * K + 31) / 32) * 32;
    const size_t smem = (size_t)PTS * 2 * D * sizeof(double) + (size_t)PTS * sizeof(T);
    const int64_t grid = (N + PTS - 1) / PTS;
    local_step_bwd_kernel<T, DT, TH><<<(unsigned)grid, threads, smem, st>>>(N, K, D, S, PTS, den_mode, eta1, eta2d, phi_rec,
                                                                        theta_rec, noise, seed, log_r, gx, glr, greg,
                                                                        greg_dev, eta1_bar, eta2d_bar, kacc);
    return cudaGetLastError();
}

template <typename T>
static int svae_local_step_bwd(int64_t N, int K, int D, int S, const T* eta1, const T* eta2d, const T* eta1_phi2,
                               const T* L_raw, const T* pi_raw, const T* phi_rec, const T* theta_rec, int den_mode,
                               const T* noise, uint64_t seed, const T* log_r, const T* gx, const T* glr, double greg,
                               const T* greg_dev, T* eta1_bar, T* eta2d_bar, T* h2_bar, T* L_raw_bar, T* pi_raw_bar,
                               T* theta_rec_bar, void* work, size_t work_bytes, cudaStream_t st) {
    if (N < 0 || K < 1 || S < 1) return VMP_E_BADARG;
    if (D < 1 || D > VMP_MAX_D) return VMP_E_BADDIM;
    const bool block_path = D > BWD_MAX_D || K > 256;            // matrices in shared memory, one CTA per point
    if (block_path && bwb_smem<T>(K, D, theta_rec_bar != nullptr) > 220 * 1024) return VMP_E_BADARG;
    if (den_mode != VMP_DEN_GAUSS && den_mode != VMP_DEN_STUDENT) return VMP_E_BADMODE;
    const size_t kl = (size_t)bwd_klen(D);
    if (!work || work_bytes < (size_t)K * kl * sizeof(double)) return VMP_E_BADARG;
    if (!eta1_phi2 || !L_raw || !pi_raw || !h2_bar || !L_raw_bar || !pi_raw_bar) return VMP_E_BADARG;
    double* kacc = static_cast<double*>(work);
    { cudaError_t me = cudaMemsetAsync(kacc, 0, (size_t)K * kl * sizeof(double), st); if (me != cudaSuccess) return (int)me; }
    if (N > 0) {
        if (!eta1 || !eta2d || !phi_rec || !theta_rec || !log_r || !gx || !glr || !eta1_bar || !eta2d_bar)
            return VMP_E_BADARG;
        cudaError_t e;
        if (block_path) {
            const size_t sm = bwb_smem<T>(K, D, theta_rec_bar != nullptr);
            if (N > 0x7fffffffLL) return VMP_E_BADARG;
            if (theta_rec_bar) {
                auto kern = local_step_bwd_block_kernel<T, true>;
                e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
                if (e != cudaSuccess) return (int)e;
                kern<<<(unsigned)N, BWB_THREADS, sm, st>>>(N, K, D, S, den_mode, eta1, eta2d, phi_rec, theta_rec, noise, seed, log_r,
                                                          gx, glr, (T)greg, greg_dev, eta1_bar, eta2d_bar, kacc);
            } else {
                auto kern = local_step_bwd_block_kernel<T, false>;
                e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
                if (e != cudaSuccess) return (int)e;
                kern<<<(unsigned)N, BWB_THREADS, sm, st>>>(N, K, D, S, den_mode, eta1, eta2d, phi_rec, theta_rec, noise, seed, log_r,
                                                          gx, glr, (T)greg, greg_dev, eta1_bar, eta2d_bar, kacc);
            }
            e = cudaGetLastError();
        } else {
#define VMP_BWD_CASE(DD)                                                                                              \
    case DD:                                                                                                          \
        e = theta_rec_bar ? launch_bwd_main<T, DD, true>(N, K, D, S, den_mode, eta1, eta2d, phi_rec, theta_rec, noise,   \
                                                         seed, log_r, gx, glr, (T)greg, greg_dev, eta1_bar, eta2d_bar, kacc, st)  \
                          : launch_bwd_main<T, DD, false>(N, K, D, S, den_mode, eta1, eta2d, phi_rec, theta_rec, noise,  \
                                                          seed, log_r, gx, glr, (T)greg, greg_dev, eta1_bar, eta2d_bar, kacc, st); \
        break;
        switch (D) {
            VMP_BWD_CASE(1) VMP_BWD_CASE(2) VMP_BWD_CASE(3) VMP_BWD_CASE(4) VMP_BWD_CASE(5) VMP_BWD_CASE(6)
            VMP_BWD_CASE(7) VMP_BWD_CASE(8)
            default:
                e = theta_rec_bar ? launch_bwd_main<T, 0, true>(N, K, D, S, den_mode, eta1, eta2d, phi_rec, theta_rec, noise,
                                                                seed, log_r, gx, glr, (T)greg, greg_dev, eta1_bar, eta2d_bar, kacc, st)
                                  : launch_bwd_main<T, 0, false>(N, K, D, S, den_mode, eta1, eta2d, phi_rec, theta_rec, noise,
                                                                 seed, log_r, gx, glr, (T)greg, greg_dev, eta1_bar, eta2d_bar,
                                                                 kacc, st);
        }
#undef VMP_BWD_CASE
        }
        if (e != cudaSuccess) return (int)e;
    }
    const size_t esm = (size_t)(4 * D * (D + 1) + 3 * D + 32) * sizeof(double);
    if (esm + 256 > 48 * 1024) {
        cudaError_t ee = cudaFuncSetAttribute(local_step_bwd_epilogue_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)esm);
        if (ee != cudaSuccess) return (int)ee;
    }
    local_step_bwd_epilogue_kernel<T><<<K, 128, esm, st>>>(K, D, eta1_phi2, L_raw, pi_raw, kacc, h2_bar, L_raw_bar,
                                                           pi_raw_bar, theta_rec_bar);
    return launch_status();
}

}  // namespace vmp

extern "C" {
size_t vmp_svae_local_step_bwd_workspace_bytes(int K, int D) {
    return (size_t)K * (size_t)vmp::bwd_klen(D) * sizeof(double);
}
int vmp_svae_local_step_bwd_f32(int64_t N, int K, int D, int S, const float* eta1, const float* eta2_diag,
                                const float* eta1_phi2, const float* L_raw, const float* pi_raw, const float* phi_rec,
                                const float* theta_rec, int den_mode, const float* noise, uint64_t seed,
                                const float* log_r, const float* gx, const float* glr, double greg, const float* greg_dev,
                                float* eta1_bar,
                                float* eta2_diag_bar, float* eta1_phi2_bar, float* L_raw_bar, float* pi_raw_bar,
                                float* theta_rec_bar, void* workspace, size_t workspace_bytes, void* stream) {
    return vmp::svae_local_step_bwd<float>(N, K, D, S, eta1, eta2_diag, eta1_phi2, L_raw, pi_raw, phi_rec, theta_rec,
                                           den_mode, noise, seed, log_r, gx, glr, greg, greg_dev, eta1_bar, eta2_diag_bar,
                                           eta1_phi2_bar, L_raw_bar, pi_raw_bar, theta_rec_bar, workspace, workspace_bytes,
                                           static_cast<cudaStream_t>(stream));
}
int vmp_svae_local_step_bwd_f64(int64_t N, int K, int D, int S, const double* eta1, const double* eta2_diag,
                                const double* eta1_phi2, const double* L_raw, const double* pi_raw,
                                const double* phi_rec, const double* theta_rec, int den_mode, const double* noise,
                                uint64_t seed, const double* log_r, const double* gx, const double* glr, double greg,
                                const double* greg_dev, double* eta1_bar, double* eta2_diag_bar, double* eta1_phi2_bar, double* L_raw_bar,
                                double* pi_raw_bar, double* theta_rec_bar, void* workspace, size_t workspace_bytes, void* stream) {
    return vmp::svae_local_step_bwd<double>(N, K, D, S, eta1, eta2_diag, eta1_phi2, L_raw, pi_raw, phi_rec, theta_rec,
                                            den_mode, noise, seed, log_r, gx, glr, greg, greg_dev, eta1_bar, eta2_diag_bar,
                                            eta1_phi2_bar, L_raw_bar, pi_raw_bar, theta_rec_bar, workspace, workspace_bytes,
                                            static_cast<cudaStream_t>(stream));
}
}
