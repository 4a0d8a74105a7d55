// block_linalg.cuh — CTA-cooperative dense helpers on double matrices in shared memory (K-sized prologues).
#pragma once
#include "common.cuh"

namespace vmp {

// In-place lower Cholesky of the D x D matrix A (leading dimension ld) in shared memory, all threads
// of the CTA cooperate.  Upper triangle is left untouched.  Returns nothing; a non-positive pivot gives NaN.
__device__ inline void chol_lower_block(double* A, int D, int ld) {
    for (int j = 0; j < D; ++j) {
        __syncthreads();
        const double djj = sqrt(A[j * ld + j]);
        __syncthreads();
        if (threadIdx.x == 0) A[j * ld + j] = djj;
        const double inv = 1.0 / djj;
        for (int i = j + 1 + threadIdx.x; i < D; i += blockDim.x) A[i * ld + j] *= inv;
        __syncthreads();
        // trailing update: A[i][c] -= A[i][j] * A[c][j], j < c <= i
        const int m = D - j - 1;
        for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
            const int i = j + 1 + e / m, c = j + 1 + e % m;
            if (c <= i) A[i * ld + c] -= A[i * ld + j] * A[c * ld + j];
        }
    }
    __syncthreads();
}

// W = L^-1 for lower-triangular L (both D x D, ld); thread t solves column t.  W's upper triangle is zeroed.
__device__ inline void tri_inverse_block(const double* L, double* W, int D, int ld) {
    for (int j = threadIdx.x; j < D; j += blockDim.x) {
        for (int i = 0; i < j; ++i) W[i * ld + j] = 0.0;
        W[j * ld + j] = 1.0 / L[j * ld + j];
        for (int i = j + 1; i < D; ++i) {
            double s = 0.0;
            for (int c = j; c < i; ++c) s += L[i * ld + c] * W[c * ld + j];
            W[i * ld + j] = -s / L[i * ld + i];
        }
    }
    __syncthreads();
}

}  // namespace vmp
