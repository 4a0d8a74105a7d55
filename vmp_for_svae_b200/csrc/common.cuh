// common.cuh — shared device helpers for libvmp_svae (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/vmp_svae.h"

#define VMP_LOG_2PI 1.8378770664093454835606594728112
#define VMP_LOG_PI 1.1447298858494001741434273513531
#define VMP_LOG_2 0.69314718055994530941723212145818

namespace vmp {

__host__ __device__ inline int phi_record_len(int D) { return D * D + 2 * D + 4; }
__host__ __device__ inline int theta_record_len(int D) { return D * D + D + 4; }
__host__ __device__ inline int stats_len(int D) { return D * D + D + 2; }

inline int launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? VMP_OK : (int)e;
}

// ---- scalar math dispatch (float keeps the accurate libdevice versions: parity first) ----------------
template <typename T> __device__ __forceinline__ T t_exp(T x);
template <> __device__ __forceinline__ float t_exp<float>(float x) { return expf(x); }
template <> __device__ __forceinline__ double t_exp<double>(double x) { return exp(x); }
template <typename T> __device__ __forceinline__ T t_log(T x);
template <> __device__ __forceinline__ float t_log<float>(float x) { return logf(x); }
template <> __device__ __forceinline__ double t_log<double>(double x) { return log(x); }
template <typename T> __device__ __forceinline__ T t_log1p(T x);
template <> __device__ __forceinline__ float t_log1p<float>(float x) { return log1pf(x); }
template <> __device__ __forceinline__ double t_log1p<double>(double x) { return log1p(x); }
template <typename T> __device__ __forceinline__ T t_sqrt(T x);
template <> __device__ __forceinline__ float t_sqrt<float>(float x) { return sqrtf(x); }
template <> __device__ __forceinline__ double t_sqrt<double>(double x) { return sqrt(x); }
template <typename T> __device__ __forceinline__ T t_softplus(T x) {
    // log(1 + e^x), stable on both tails (tf.nn.softplus)
    return x > T(0) ? x + t_log1p(t_exp(-x)) : t_log1p(t_exp(x));
}

// digamma for x > 0 (recurrence to x >= 6, then the asymptotic series); double only, K-sized callers.
__device__ inline double digamma_pos(double x) {
    double r = 0.0;
    while (x < 6.0) { r -= 1.0 / x; x += 1.0; }
    const double f = 1.0 / (x * x);
    // psi(x) ~ ln x - 1/(2x) - 1/(12x^2) + 1/(120x^4) - 1/(252x^6) + 1/(240x^8) - 5/(660x^10) + 691/(32760x^12) - 1/(12x^14)
    const double t = f * (-1.0 / 12.0 + f * (1.0 / 120.0 + f * (-1.0 / 252.0 + f * (1.0 / 240.0 +
                     f * (-5.0 / 660.0 + f * (691.0 / 32760.0 + f * (-1.0 / 12.0)))))));
    return r + log(x) - 0.5 / x + t;
}

// ---- warp helpers -------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <typename T> __device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// block-wide sum of a double (blockDim.x <= 1024); result valid in thread 0
__device__ inline double block_sum(double v, double* smem32) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) smem32[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        r = lane < nw ? smem32[lane] : 0.0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}

// ---- Philox4x32-10 counter-based generator (Salmon et al. 2011) -------------------------------------
struct Philox {
    __device__ static __forceinline__ uint4 gen(uint4 ctr, uint2 key) {
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
            ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
            key.x += 0x9E3779B9u;
            key.y += 0xBB67AE85u;
        }
        return ctr;
    }
};

// uniform in (0,1): 23 random bits at the bin centres (k + 1/2) / 2^23 — every value is exactly representable in fp32
// (k + 1/2 needs 24 significant bits), so the result is never 0 or 1 (safe for log / Box-Muller / Gumbel)
__host__ __device__ __forceinline__ float u32_to_unit(uint32_t x) { return ((float)(x >> 9) + 0.5f) * (1.0f / 8388608.0f); }

// the 9 low bits of three Philox words that u32_to_unit leaves unused -> one more uniform (27 bits, the top 23 used)
__host__ __device__ __forceinline__ float philox_spare_uniform(uint4 r) {
    return u32_to_unit(((r.x & 0x1ffu) << 23) | ((r.y & 0x1ffu) << 14) | ((r.z & 0x1ffu) << 5));
}
// four standard normals for (pair, sample s, dims 4q..4q+3) — the definition of the in-kernel noise stream; *spare = the
// uniform made of the block's unused low bits (the Gumbel uniform of the pair when s = q = 0, see philox_uniform_pair)
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint64_t pair, uint32_t s, uint32_t q, float* spare = nullptr) {
    const uint4 r = Philox::gen(make_uint4((uint32_t)pair, (uint32_t)(pair >> 32), s, q),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    if (spare != nullptr) *spare = philox_spare_uniform(r);
    const float u0 = u32_to_unit(r.x), u1 = u32_to_unit(r.y), u2 = u32_to_unit(r.z), u3 = u32_to_unit(r.w);
    // Box-Muller on the special-function unit: radius sqrt(-2 ln u) from MUFU.LG2 + MUFU.SQRT, direction from MUFU.SIN / MUFU.COS
    // at theta = 2 pi u - pi in (-pi, pi) (absolute error ~5e-7): 12 instructions per pair of normals instead of ~60 for
    // logf / sqrtf / sincospif — the noise generation was 8 % of the De=32 engine's instructions.  This IS the definition of the
    // in-kernel stream (vmp_fill_noise writes the same values); parity against the oracle always injects the noise.
    float l0, l1, r0, r1, s0, c0, s1, c1;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l0) : "f"(u0));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l1) : "f"(u2));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(-1.3862943611198906f * l0));      // -2 ln 2 * log2 u
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(-1.3862943611198906f * l1));
    const float t0 = fmaf(6.283185307179586f, u1, -3.141592653589793f), t1 = fmaf(6.283185307179586f, u3, -3.141592653589793f);
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s0) : "f"(t0));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c0) : "f"(t0));
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s1) : "f"(t1));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c1) : "f"(t1));
    return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}
__device__ __forceinline__ float philox_normal1(uint64_t seed, uint64_t pair, uint32_t s, uint32_t d) {
    const float4 v = philox_normal4(seed, pair, s, d >> 2);
    const uint32_t c = d & 3u;
    return c == 0 ? v.x : c == 1 ? v.y : c == 2 ? v.z : v.w;
}
// the uniform behind the Gumbel-max categorical draw of pair (n,k) — the in-kernel definition of u[n,k]: the spare low bits of
// the Philox block that also yields the pair's first four normals (s = 0, q = 0), so a kernel that draws those normals gets the
// uniform without a second Philox call
__device__ __forceinline__ float philox_uniform_pair(uint64_t seed, uint64_t pair) {
    const uint4 r = Philox::gen(make_uint4((uint32_t)pair, (uint32_t)(pair >> 32), 0u, 0u),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    return philox_spare_uniform(r);
}
// Gumbel(0,1) from a uniform: tf.multinomial's GPU kernel draws z = argmax_k(logit_k - log(-log(u_k)))
template <typename T> __device__ __forceinline__ T gumbel_from_uniform(T u) { return -t_log(-t_log(u)); }

}  // namespace vmp
