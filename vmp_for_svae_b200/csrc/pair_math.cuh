// pair_math.cuh — thread-per-pair evaluation of one (point n, component k) Gaussian posterior solve.
//
// Math (verified against the literal reference graph in tests/):
//   P1 = diag(p1), p1 = -2 eta2_diag;  mu1 = eta1 / p1;  P~ = P2_k + P1 = L L^T;  d = mu1 - mu2_k
//   a  = L^-1 (P2_k d),  a1 = L^-1 (P1 d)
//   score_nk = log pi_k - 1/2 a.a1 + 1/2 logdet P2_k - sum_i log L_ii     (per-point terms cancel in the softmax)
//   x_nks    = mu1 + L^-T (eps_nks - a)                                    (= P~^-1 eta~1 + L^-T eps, svae.py:111-118)
//   log N(x_nks | phi~_nk) = -1/2 |eps|^2 + sum_i log L_ii - D/2 log 2pi   (gaussian.py:74-105 on these samples)
//   Delta^2 = |W_k (x - m_k)|^2 with the lower-triangular W_k of the theta record (prepare.cu)
//
// DT > 0: D is a compile-time constant, every loop unrolls and L lives in registers (D <= 8).
// DT == 0: run-time D <= VMP_MAX_D, arrays live in local memory (generic fallback).
#pragma once
#include "common.cuh"

namespace vmp {

template <typename T, int DT>
struct PairMath {
    static constexpr int DM = DT ? DT : VMP_MAX_D;
    static constexpr int LM = DM * (DM + 1) / 2;

    T L[LM];     // packed lower Cholesky factor, row-major: (i,j) at i(i+1)/2 + j
    T a[DM];     // L^-1 P2 d
    T mu1[DM];
    T hld;       // sum_i log L_ii = 1/2 logdet P~
    T score;     // unnormalised log responsibility
    int bad;     // non-positive pivot seen

    __device__ __forceinline__ static int tri(int i, int j) { return i * (i + 1) / 2 + j; }

    // phi record of component k: P2[D*D] | mu2[D] | h2[D] | log_pi, logdetP2
    __device__ __forceinline__ void factor(const int Drt, const T* __restrict__ eta1_n,
                                           const T* __restrict__ eta2d_n, const T* __restrict__ prec) {
        const int D = DT ? DT : Drt;
        const T* P2 = prec;
        const T* mu2 = prec + D * D;
        T d[DM], g[DM], g1[DM], p1[DM];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            p1[i] = T(-2) * eta2d_n[i];
            mu1[i] = eta1_n[i] / p1[i];
            d[i] = mu1[i] - mu2[i];
            g1[i] = p1[i] * d[i];
        }
#pragma unroll
        for (int i = 0; i < D; ++i) {
            T s = T(0);
#pragma unroll
            for (int c = 0; c < D; ++c) s = fma(P2[i * D + c], d[c], s);
            g[i] = s;
        }
        bad = 0;
        hld = T(0);
        T quad = T(0);
        T a1[DM];
        // Cholesky-Banachiewicz row by row, forward substitution of both right-hand sides fused in
#pragma unroll
        for (int i = 0; i < D; ++i) {
            T sa = g[i], sa1 = g1[i];
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                T s = P2[i * D + j];
                if (j == i) s += p1[i];
#pragma unroll
                for (int c = 0; c < j; ++c) s = fma(-L[tri(i, c)], L[tri(j, c)], s);
                if (j == i) {
                    if (!(s > T(0))) bad = 1;
                    const T l = t_sqrt(s);
                    L[tri(i, i)] = l;
                    hld += t_log(l);
                    a[i] = sa / l;
                    a1[i] = sa1 / l;
                    quad = fma(a[i], a1[i], quad);
                } else {
                    const T l = s / L[tri(j, j)];
                    L[tri(i, j)] = l;
                    sa = fma(-l, a[j], sa);
                    sa1 = fma(-l, a1[j], sa1);
                }
            }
        }
        const T log_pi = prec[D * D + 2 * D], logdetP2 = prec[D * D + 2 * D + 1];
        score = log_pi - T(0.5) * quad + T(0.5) * logdetP2 - hld;
    }

    // x = mu1 + L^-T (eps - a)
    __device__ __forceinline__ void sample(const int Drt, const T* eps, T* x) const {
        const int D = DT ? DT : Drt;
#pragma unroll
        for (int ii = 0; ii < D; ++ii) {
            const int i = D - 1 - ii;
            T s = eps[i] - a[i];
#pragma unroll
            for (int c = i + 1; c < D; ++c) s = fma(-L[tri(c, i)], x[c], s);
            x[i] = s / L[tri(i, i)];
        }
#pragma unroll
        for (int i = 0; i < D; ++i) x[i] += mu1[i];
    }

    // inverse of sample(): eps = a + L^T (x - mu1)  (used when the caller supplies the samples, svae.compute_elbo)
    __device__ __forceinline__ void eps_from_x(const int Drt, const T* x, T* eps) const {
        const int D = DT ? DT : Drt;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            T s = a[i];
#pragma unroll
            for (int c = i; c < D; ++c) s = fma(L[tri(c, i)], x[c] - mu1[c], s);
            eps[i] = s;
        }
    }

    // Delta^2 = |W (x - m)|^2, theta record: W[D*D] (lower) | m[D] | cden, nu, elogpi, logdetP
    __device__ __forceinline__ static T maha(const int Drt, const T* __restrict__ trec, const T* x) {
        const int D = DT ? DT : Drt;
        const T* W = trec;
        const T* m = trec + D * D;
        T xm[DM];
#pragma unroll
        for (int i = 0; i < D; ++i) xm[i] = x[i] - m[i];
        T q2 = T(0);
#pragma unroll
        for (int i = 0; i < D; ++i) {
            T s = T(0);
#pragma unroll
            for (int c = 0; c <= i; ++c) s = fma(W[i * D + c], xm[c], s);
            q2 = fma(s, s, q2);
        }
        return q2;
    }
};

// log-density of the ELBO denominator given Delta^2 (constants live in the theta record)
template <typename T>
__device__ __forceinline__ T den_logprob(int den_mode, int D, T maha, T cden, T nu) {
    if (den_mode == VMP_DEN_GAUSS) return cden - T(0.5) * maha;                   // gaussian.py:74-105 at E[theta]
    return cden - T(0.5) * (nu + T(D)) * t_log1p(maha / nu);                      // student_t.py:39
}

}  // namespace vmp
