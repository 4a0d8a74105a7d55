// suffstats_mma.cu — responsibility-weighted sufficient statistics for D in {16, 32} on the tensor cores (svae.m_step,
// /root/reference/models/svae.py:154-176: N_k = sum_n r_nk, sum_n r_nk x_n, sum_n r_nk x_n x_n^T; with u_nk the SMM weights
// w = r u of smm.m_step, /root/reference/models/smm.py:25-85).
//
// Per component the D x D statistic is X^T diag(w_k) X: a GEMM whose contraction index is the point.  Warp-level
// mma.sync.m16n8k8 (tf32 operands split hi/lo, hi*hi + lo*hi + hi*lo: fp32-accurate products; the tensor core only ever sums the
// 8 points of one step, the running sums are kept with round-to-nearest FADDs because its fp32 accumulation truncates).
//   A fragment = (w_k x)^T: row = coordinate i, column = point;  B fragment = x: row = point, column = coordinate j —
//   both are built from the SAME eight x values a lane holds (x[t][g + 8 q], x[t + 4][g + 8 q]), so a group of 8 points costs
//   8 shared-memory loads per warp plus, per component, 2 weight loads, D/4 multiplies and the splits.
// Only the tiles of the lower block triangle are multiplied (6 of 8 at D = 32) and only entries i >= j are written (mirrored),
// so the statistic is exactly symmetric.  sum w x and sum w / sum r ride along as FADDs on the A-fragment values.
// A CTA = 8 warps x 2 components = 16 components over one slice of the points; chunks of 64 points (x rows + the 16 weight
// columns) stream through a double-buffered cp.async stage shared by the 8 warps.  The natural-gradient update runs in the
// tail of the last CTA (ng_tail.cuh) as in every other statistics kernel.
#include "common.cuh"
#include "ng_tail.cuh"

namespace vmp {

namespace {
constexpr int SM_WARPS = 8, SM_CPW = 2, SM_KC = SM_WARPS * SM_CPW, SM_CH = 64;

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_tf32_zero(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(v) & 0xffffe000u;
    lo = __float_as_uint(v - __uint_as_float(hi));
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    const int sz = valid ? 16 : 0;                                   // src-size 0: zero-fill, nothing is read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}

template <int D> struct MmaGeom {
    static constexpr int MT = D / 16, NT = D / 8;                    // m-tiles (16 rows), n-tiles (8 columns)
    static constexpr int XS = D + 8;                                 // x row stride: the (t, g) fragment reads hit 32 banks
    // tiles (mt, nt) of the lower block triangle: nt <= 2 mt + 1
    static constexpr int NTILES = MT * (MT + 1);                     // sum_mt (2 mt + 2)
    __host__ __device__ static constexpr int tile_index(int mt, int nt) { return mt * (mt + 1) + nt; }
};

template <int D>
__global__ void __launch_bounds__(SM_WARPS * 32, 2)
suffstats_mma_kernel(int64_t N, int K, int64_t pts_per_slice, const float* __restrict__ x, const float* __restrict__ r,
                     int r_is_log, const float* __restrict__ u, double* __restrict__ stats, const NgTail tail) {
    using G = MmaGeom<D>;
    constexpr int MT = G::MT, NT = G::NT, XS = G::XS, NTILES = G::NTILES;
    extern __shared__ __align__(16) float smm_raw[];
    float* xs0 = smm_raw;                                            // [2][SM_CH][XS]
    float* rs0 = xs0 + 2 * SM_CH * XS;                               // [2][SM_CH][SM_KC]
    float* us0 = rs0 + 2 * SM_CH * SM_KC;                            // [2][SM_CH][SM_KC]   (SMM only)
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5, g = lane >> 2, t = lane & 3;
    const int k0 = blockIdx.x * SM_KC, kw = k0 + wib * SM_CPW;       // first component of the CTA / of this warp
    const int64_t n_begin = (int64_t)blockIdx.y * pts_per_slice, n_end = min(N, n_begin + pts_per_slice);
    const bool has_u = u != nullptr;
    const int SL = stats_len(D);

    auto issue = [&](int64_t n0, int buf) {                          // chunk of SM_CH points starting at n0 -> stage buf
        float* xd = xs0 + buf * SM_CH * XS;
        float* rd = rs0 + buf * SM_CH * SM_KC;
        float* ud = us0 + buf * SM_CH * SM_KC;
        for (int e = tid; e < SM_CH * (D / 4); e += SM_WARPS * 32) {
            const int p = e / (D / 4), c = e - p * (D / 4);
            const bool ok = n0 + p < n_end;
            cp_async16(xd + p * XS + 4 * c, ok ? x + (n0 + p) * D + 4 * c : x, ok);
        }
        for (int e = tid; e < SM_CH * (SM_KC / 4); e += SM_WARPS * 32) {
            const int p = e / (SM_KC / 4), c = e - p * (SM_KC / 4);
            const bool ok = n0 + p < n_end && k0 + 4 * c < K;
            cp_async16(rd + p * SM_KC + 4 * c, ok ? r + (n0 + p) * K + k0 + 4 * c : r, ok);
            if (has_u) cp_async16(ud + p * SM_KC + 4 * c, ok ? u + (n0 + p) * K + k0 + 4 * c : u, ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    float acc[SM_CPW][NTILES][4];
    float sx[SM_CPW][MT][2], sw[SM_CPW], sr[SM_CPW];
#pragma unroll
    for (int c = 0; c < SM_CPW; ++c) {
#pragma unroll
        for (int q = 0; q < NTILES; ++q)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[c][q][e] = 0.f;
#pragma unroll
        for (int m = 0; m < MT; ++m) sx[c][m][0] = sx[c][m][1] = 0.f;
        sw[c] = sr[c] = 0.f;
    }
    const bool warp_on = kw < K;

    issue(n_begin, 0);
    int buf = 0;
    for (int64_t n0 = n_begin; n0 < n_end; n0 += SM_CH, buf ^= 1) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                             // chunk n0 has landed; every warp is done with stage buf^1
        if (n0 + SM_CH < n_end) issue(n0 + SM_CH, buf ^ 1);
        if (!warp_on) continue;
        const float* xb = xs0 + buf * SM_CH * XS;
        const float* rb = rs0 + buf * SM_CH * SM_KC + wib * SM_CPW;
        const float* ub = us0 + buf * SM_CH * SM_KC + wib * SM_CPW;
#pragma unroll 1
        for (int p0 = 0; p0 < SM_CH && n0 + p0 < n_end; p0 += 8) {
            // the lane's eight x values: points t, t+4, coordinates g + 8 q
            uint32_t xh[2][NT], xl[2][NT];
            float xv[2][NT];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int q = 0; q < NT; ++q) {
                    xv[h][q] = xb[(p0 + t + 4 * h) * XS + g + 8 * q];
                    split_tf32(xv[h][q], xh[h][q], xl[h][q]);
                }
            const bool in0 = n0 + p0 + t < n_end, in1 = n0 + p0 + t + 4 < n_end;
#pragma unroll
            for (int c = 0; c < SM_CPW; ++c) {
                float r0 = rb[(p0 + t) * SM_KC + c], r1 = rb[(p0 + t + 4) * SM_KC + c];
                if (r_is_log) { r0 = __expf(r0); r1 = __expf(r1); }
                r0 = (in0 && kw + c < K) ? r0 : 0.f;
                r1 = (in1 && kw + c < K) ? r1 : 0.f;
                const float w0 = has_u ? r0 * ub[(p0 + t) * SM_KC + c] : r0;
                const float w1 = has_u ? r1 * ub[(p0 + t + 4) * SM_KC + c] : r1;
                sr[c] += r0 + r1;
                sw[c] += w0 + w1;
                // A fragments of the m-tiles: rows g + 16 m, g + 8 + 16 m; columns = points t, t+4
                uint32_t ah[MT][4], al[MT][4];
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    const float a0 = w0 * xv[0][2 * m], a1 = w0 * xv[0][2 * m + 1];
                    const float a2 = w1 * xv[1][2 * m], a3 = w1 * xv[1][2 * m + 1];
                    sx[c][m][0] += a0 + a2;
                    sx[c][m][1] += a1 + a3;
                    split_tf32(a0, ah[m][0], al[m][0]);
                    split_tf32(a1, ah[m][1], al[m][1]);
                    split_tf32(a2, ah[m][2], al[m][2]);
                    split_tf32(a3, ah[m][3], al[m][3]);
                }
#pragma unroll
                for (int m = 0; m < MT; ++m)
#pragma unroll
                    for (int q = 0; q <= 2 * m + 1; ++q) {
                        float d[4];
                        mma_tf32_zero(d, al[m], xh[0][q], xh[1][q]);
                        mma_tf32(d, ah[m], xl[0][q], xl[1][q]);
                        mma_tf32(d, ah[m], xh[0][q], xh[1][q]);
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[c][G::tile_index(m, q)][e] += d[e];
                    }
            }
        }
    }
    // ---- this warp's sums -> the global fp64 statistics.  C fragment: e0/e1 = (row g + 16 m, columns 8 q + 2 t, + 1),
    // e2/e3 = (row g + 8 + 16 m, the same columns); only i >= j is written, and mirrored
    if (warp_on) {
#pragma unroll
        for (int c = 0; c < SM_CPW; ++c) {
            if (kw + c >= K) continue;
            double* out = stats + (size_t)(kw + c) * SL;
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int q = 0; q <= 2 * m + 1; ++q)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int i = g + 16 * m + 8 * (e >> 1), j = 8 * q + 2 * t + (e & 1);
                        const double v = (double)acc[c][G::tile_index(m, q)][e];
                        if (i >= j && v != 0.0) {
                            atomicAdd(out + 2 + D + i * D + j, v);
                            if (i != j) atomicAdd(out + 2 + D + j * D + i, v);
                        }
                    }
            // sum w x: the four t-lanes of a g hold partial sums of the same rows; sum w, sum r: every lane of a t holds the same
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float v = sx[c][m][h];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    if (t == 0 && v != 0.f) atomicAdd(out + 2 + g + 16 * m + 8 * h, (double)v);
                }
            float vw = sw[c], vr = sr[c];
            vw += __shfl_xor_sync(0xffffffffu, vw, 1); vw += __shfl_xor_sync(0xffffffffu, vw, 2);
            vr += __shfl_xor_sync(0xffffffffu, vr, 1); vr += __shfl_xor_sync(0xffffffffu, vr, 2);
            if (lane == 0) {
                atomicAdd(out + 0, (double)vr);
                atomicAdd(out + 1, (double)vw);
            }
        }
    }
    ng_tail_run<float>(tail, K, D, stats, gridDim.x * gridDim.y);
}

template <int D>
int launch_mma(int64_t N, int K, const float* x, const float* r, int r_is_log, const float* u, double* stats, const NgTail& tail,
               cudaStream_t st) {
    using G = MmaGeom<D>;
    const int kgroups = (K + SM_KC - 1) / SM_KC;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // two CTAs per SM; a slice is at least 256 points (its flush costs ~D^2 atomics per component) and at most 1024 (fp32 sums)
    int64_t nslices = (2 * (int64_t)sms + kgroups - 1) / kgroups;
    int64_t pps = (N + nslices - 1) / nslices;
    pps = ((pps + SM_CH - 1) / SM_CH) * SM_CH;
    if (pps < 256) pps = 256;
    if (pps > 1024) pps = 1024;
    nslices = (N + pps - 1) / pps;
    if (nslices > 65535) return -100;
    const size_t smem = sizeof(float) * (2 * (size_t)SM_CH * G::XS + 4 * (size_t)SM_CH * SM_KC);
    auto kern = suffstats_mma_kernel<D>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<dim3(kgroups, (unsigned)nslices), SM_WARPS * 32, smem, st>>>(N, K, pps, x, r, r_is_log, u, stats, tail);
    return launch_status();
}
}  // namespace

// fp32, D in {16, 32}, K % 4 == 0, 16-byte aligned x / r / u; -100 = not this kernel's shape
int suffstats_mma(int64_t N, int K, int D, const float* x, const float* r, int r_is_log, const float* u, double* stats,
                  const NgTail& tail, cudaStream_t st) {
    if ((D != 16 && D != 32) || K % 4 != 0 || N < 64) return -100;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(r) | reinterpret_cast<uintptr_t>(u)) % 16 != 0) return -100;
    return D == 32 ? launch_mma<32>(N, K, x, r, r_is_log, u, stats, tail, st) : launch_mma<16>(N, K, x, r, r_is_log, u, stats, tail, st);
}

}  // namespace vmp
