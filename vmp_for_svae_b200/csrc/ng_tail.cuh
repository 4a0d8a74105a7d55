// ng_tail.cuh — the natural-gradient (CVI) update fused into the TAIL of the sufficient-statistics reduction
// (north_star: "The NIW/Dirichlet natural-gradient step is fused into the tail of that reduction"; svae.m_step 154-176 +
// svae.update_gmm_params 376-403).  Every CTA of a statistics kernel adds its partial sums to the global fp64 statistics
// with atomics and then takes a ticket; the CTA that draws the last ticket knows that all partials have landed and applies
//     theta <- (1 - rho) theta + rho (prior + [N_k, sum r x x^T, sum r x, N_k, N_k + 1])
// for all K components — no second launch.  Used on the single-GPU path; with several ranks the statistics are all-reduced
// between the reduction and the update, which therefore stays a separate (K-sized) launch.
#pragma once
#include "common.cuh"

namespace vmp {

struct NgTail {
    unsigned int* counter;      // zero before the launch; the last CTA resets it.  nullptr: no fused update
    double rho;
    const double* rho_dev;      // device-resident step size (overrides rho) or nullptr
    int only_alpha;             // svae.m_step_smm: only the Dirichlet parameter moves
    const void *p_alpha, *p_A, *p_b, *p_beta, *p_vhat;
    void *alpha, *A, *b, *beta, *v_hat;
};

inline NgTail ng_tail_none() {
    NgTail t;
    t.counter = nullptr; t.rho = 0.0; t.rho_dev = nullptr; t.only_alpha = 0;
    t.p_alpha = t.p_A = t.p_b = t.p_beta = t.p_vhat = nullptr;
    t.alpha = t.A = t.b = t.beta = t.v_hat = nullptr;
    return t;
}

// Called by ALL threads of EVERY CTA after the CTA's own atomics; total_ctas = number of CTAs of the launch.
template <typename T>
__device__ __forceinline__ void ng_tail_run(const NgTail& t, int K, int D, const double* __restrict__ stats,
                                            unsigned int total_ctas) {
    if (t.counter == nullptr) return;
    __shared__ unsigned int s_ticket;
    __threadfence();                                     // this thread's atomics are ordered before the ticket
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) s_ticket = atomicAdd(t.counter, 1u);
    __syncthreads();
    if (s_ticket != total_ctas - 1) return;
    __threadfence();
    const double rho = t.rho_dev != nullptr ? __ldcg(t.rho_dev) : t.rho;
    const int SL = stats_len(D);
    const T* p_alpha = static_cast<const T*>(t.p_alpha); const T* p_A = static_cast<const T*>(t.p_A);
    const T* p_b = static_cast<const T*>(t.p_b); const T* p_beta = static_cast<const T*>(t.p_beta);
    const T* p_vhat = static_cast<const T*>(t.p_vhat);
    T* alpha = static_cast<T*>(t.alpha); T* A = static_cast<T*>(t.A); T* b = static_cast<T*>(t.b);
    T* beta = static_cast<T*>(t.beta); T* v_hat = static_cast<T*>(t.v_hat);
    const int tid = threadIdx.x + threadIdx.y * blockDim.x, nth = blockDim.x * blockDim.y;
    for (int k = tid; k < K; k += nth) {
        const double Nk = __ldcg(stats + (size_t)k * SL);
        alpha[k] = (T)((1.0 - rho) * (double)alpha[k] + rho * ((double)p_alpha[k] + Nk));
        if (!t.only_alpha) {
            beta[k] = (T)((1.0 - rho) * (double)beta[k] + rho * ((double)p_beta[k] + Nk));
            v_hat[k] = (T)((1.0 - rho) * (double)v_hat[k] + rho * ((double)p_vhat[k] + Nk + 1.0));
        }
    }
    if (!t.only_alpha) {
        for (int e = tid; e < K * D; e += nth) {
            const int k = e / D, i = e - k * D;
            b[e] = (T)((1.0 - rho) * (double)b[e] + rho * ((double)p_b[e] + __ldcg(stats + (size_t)k * SL + 2 + i)));
        }
        for (int e = tid; e < K * D * D; e += nth) {
            const int k = e / (D * D), o = e - k * D * D;
            A[e] = (T)((1.0 - rho) * (double)A[e] + rho * ((double)p_A[e] + __ldcg(stats + (size_t)k * SL + 2 + D + o)));
        }
    }
    if (tid == 0) *t.counter = 0u;                       // ready for the next launch
}

}  // namespace vmp
