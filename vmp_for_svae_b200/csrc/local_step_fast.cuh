// local_step_fast.cuh — the register-resident group engine of the fused local step (fp32; engine dimension D in
// {16, 32, 64}; a caller dimension 9 <= Dr <= 64 runs in the next engine size with an identity block in rows Dr..D-1).
//
// Mapping.  BS = D/4 lanes of a warp ("group") own one (point n, component k) pair; lane gl owns rows r*BS + gl,
// r < ROWS = 4, of the D x D system.  The lower triangle of P~ = P2_k + diag(p1_n) lives in registers
// (A[r][c], c < (r+1)*BS; entries right of the diagonal are don't-care slots that are computed but never feed a
// meaningful value), so the whole factorisation runs out of the register file:
//   * right-looking factorisation in the square-root-free form P~ = Lu diag(d) Lu^T (L = Lu diag(sqrt d)): at step j
//     the pivot d_j is broadcast with one shuffle, every lane publishes its RAW entries of column j to a 2 x D
//     shared-memory column buffer, scales them by 1/d_j (its multipliers Lu[.][j]) and updates its rows with the raw
//     column read back as 128-bit broadcasts.  The update is issued as packed FFMA2 (PTX fma.rn.f32x2, sm_100): two FMAs per
//     instruction with the row's multiplier as broadcast scalar operand — measured 1.65x over scalar FFMA in THIS
//     kernel because it halves the issue slots and the instruction footprint (the scalar build was instruction-fetch
//     bound; in a pure register loop both forms reach the pipe's peak, tools/fp32_peak.py);
//   * the two forward substitutions a = L^-1 P2 d, a1 = L^-1 P1 d ride along as a packed right-hand-side pair;
//   * back substitution L^T y = eps - a goes block by block (BS x BS) with the blocks transposed through a padded
//     shared-memory tile, the solved block broadcast through shared memory;
//   * the theta quadratic form |W_k (x - m_k)|^2 is a row-owned triangular mat-vec against the staged W_k.
// A CTA of WARPS warps holds PPC = WARPS*32/BS points and walks k = 0..K-1 with all its pairs on the same k, so the
// per-component record (P2_k | W_k | mu2_k | m_k | scalars, rows padded to D+4 floats: conflict-free 128-bit row
// reads) is staged ONCE per CTA per k — by a 1-D bulk TMA copy (cp.async.bulk + mbarrier complete_tx) into a
// double-buffered stage, issued one component ahead.  Per point, the K scores / ELBO terms sit in shared memory until
// the log-sum-exp; the categorical draw is an online Gumbel-max (tf.multinomial's GPU algorithm), so the selected
// sample x[n, z_n, 0] is simply the running arg-max's sample and no second pass is needed.
// Every loop whose bounds or register indices depend on an outer index is a compile-time static_for: the kernel
// is straight-line code (~8k instructions per component at D=64) with the matrix held in ~160 named registers.
#pragma once
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace vmp {

struct FastParams {
    int64_t N;
    int K, S, den_mode;
    int Dr;                 // caller's latent dimension (<= the engine's D; rows Dr..D-1 are an identity block)
    uint64_t pair_offset;   // n_offset * K: global pair index of (point 0, component 0) — key of the in-kernel noise
    const float* eta1;
    const float* eta2d;
    const float* recs;      // [K][REC] packed staged records
    const float* noise;     // [N,K,D,S] or nullptr
    const float* gum_u;     // [N,K] or nullptr
    uint64_t seed;
    float* log_r;
    float* x_sample;
    int32_t* z;
    float* x_k_samples;
    double* elbo_acc;
    int64_t ntiles;
};

template <int D, int BS_> struct FastGeom {
    static constexpr int BS = BS_;                   // lanes per pair
    static constexpr int ROWS = D / BS_;             // rows per lane
    static constexpr int LD = D + 4;                 // padded row stride of the staged matrices
    static constexpr int REC = 2 * D * LD + 2 * D + 8;
    static constexpr int TS = BS + 4;                // transpose tile stride
    // Lane numbering inside a warp (profiles/r2_smem_lane_bits.md): a 128-bit shared-memory load costs 4 wavefronts when the
    // four lanes of an aligned quad read four different chunks (address selected by lane bits b0 AND b1) and 2 otherwise.
    // The staged P2 / W rows are selected by gl, the column broadcasts by the pair; so gl must not occupy both b0 and b1, and
    // neither may the pair index.  GLMASK = the lane-id bits that hold gl (low to high); the remaining bits hold the pair.
    static constexpr unsigned GLMASK = BS_ == 16 ? 0x1eu : BS_ == 8 ? 0x0du : 0x05u;
    // group scratch: col[2][D] | vec[D] | ybf[max(BS,4)] | tbf[BS][TS] | mub[D] | p1b[D] | xb[D] | ab[D] | ib[D]; kept congruent to BS
    // mod 32 so the groups of one warp land in different banks.  mub / p1b / xb hold per-point state (mu1, p1, the
    // running arg-max sample) that would otherwise pin 12 registers across the whole factorisation.
    static constexpr int GS_RAW = 8 * D + (BS < 4 ? 4 : BS) + BS * TS;
    static constexpr int GS = ((GS_RAW - BS + 31) / 32) * 32 + BS;
};

#ifndef VMP_D32_LANES
#define VMP_D32_LANES 4
#endif

template <int D, int BS> struct FastLaunch;
template <> struct FastLaunch<64, 16> { static constexpr int WARPS = 8, MINB = 1; };
template <> struct FastLaunch<32, 8> { static constexpr int WARPS = 8, MINB = 2; };
template <> struct FastLaunch<32, 4> { static constexpr int WARPS = 8, MINB = 1; };     // 8 rows per lane, 8 pairs per warp
template <> struct FastLaunch<16, 4> { static constexpr int WARPS = 8, MINB = 2; };

__host__ __device__ inline int fast_rec_len(int D) { return 2 * D * (D + 4) + 2 * D + 8; }

template <int D, int BS>
inline size_t fast_smem_bytes(int K) {
    using G = FastGeom<D, BS>;
    constexpr int PPC = FastLaunch<D, BS>::WARPS * (32 / G::BS);
    return sizeof(float) * (2 * (size_t)G::REC + (size_t)PPC * G::GS + (size_t)PPC * K * 3);
}

// defined in fast_d16.cu / fast_d32.cu / fast_d64.cu (one translation unit per D so they compile in parallel)
template <int D, int BS> int launch_fast(const FastParams& p, cudaStream_t st);
// (phi_rec, theta_rec) of latent dimension Dr -> staged records of the engine dimension De >= Dr
void launch_pack_fast_records(int K, int Dr, int De, const float* phi_rec, const float* theta_rec, float* out,
                              cudaStream_t st);

#ifdef VMP_FAST_IMPL
// ---- mbarrier / bulk-copy primitives (inline PTX) ---------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// compile-time loops: the index is a template constant, so every inner bound / register index that depends on it is
// a constant at instantiation time (a plain `#pragma unroll` nest leaves j-dependent inner loops rolled and pushes
// the register-resident matrix into local memory)
template <int I, int N, typename F> __device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}
template <int I, typename F> __device__ __forceinline__ void static_for_down(F&& f) {   // I-1, ..., 0
    if constexpr (I > 0) {
        f(std::integral_constant<int, I - 1>{});
        static_for_down<I - 1>(f);
    }
}

// packed FP32x2 FMA of sm_100 (SASS FFMA2): two FMAs per issue slot.
// (d0, d1) += a * (b0, b1), a broadcast scalar (ptxas folds the duplicate into the scalar-operand form)
__device__ __forceinline__ void ffma2_bcast(float& d0, float& d1, float a, float b0, float b1) {
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2,%2};\n\tmov.b64 rb, {%3,%4};\n\tmov.b64 rc, {%0,%1};\n\t"
        "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0,%1}, rc;\n\t}"
        : "+f"(d0), "+f"(d1)
        : "f"(a), "f"(b0), "f"(b1));
}
// (d0, d1) += (a0, a1) * (b0, b1)
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2,%3};\n\tmov.b64 rb, {%4,%5};\n\tmov.b64 rc, {%0,%1};\n\t"
        "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0,%1}, rc;\n\t}"
        : "+f"(d0), "+f"(d1)
        : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

__device__ __forceinline__ float lg2_approx(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// scatter the low bits of v to the set bits of mask / gather them back (compile-time when v is a constant)
__host__ __device__ constexpr unsigned bits_deposit(unsigned v, unsigned mask) {
    unsigned r = 0;
    for (unsigned b = 0, k = 0; b < 32; ++b)
        if (mask >> b & 1u) r |= ((v >> k++) & 1u) << b;
    return r;
}
__host__ __device__ constexpr unsigned bits_extract(unsigned v, unsigned mask) {
    unsigned r = 0;
    for (unsigned b = 0, k = 0; b < 32; ++b)
        if (mask >> b & 1u) r |= ((v >> b) & 1u) << k++;
    return r;
}
// reductions over the lanes of one pair: butterflies over the lane bits that hold gl
template <unsigned GLMASK> __device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int b = 4; b >= 0; --b)
        if (GLMASK >> b & 1u) v += __shfl_xor_sync(0xffffffffu, v, 1 << b);
    return v;
}
template <unsigned GLMASK> __device__ __forceinline__ double group_sum_d(double v) {
#pragma unroll
    for (int b = 4; b >= 0; --b)
        if (GLMASK >> b & 1u) v += __shfl_xor_sync(0xffffffffu, v, 1 << b);
    return v;
}
// v[e] summed over the BS lanes of a pair, element e delivered to group-lane e (recursive halving: BS - 1 shuffles)
template <int BS, unsigned GLMASK> __device__ __forceinline__ float group_reduce_scatter(float (&v)[BS], int gl) {
    int n = BS;
#pragma unroll
    for (int bit = 31 - __builtin_clz((unsigned)BS) - 1; bit >= 0; --bit) {
        const int half = n / 2;
        const bool hb = (gl >> bit) & 1;
        const int lanebit = (int)bits_deposit(1u << bit, GLMASK);
#pragma unroll
        for (int e = 0; e < BS / 2; ++e) {
            if (e < half) {
                const float send = hb ? v[e] : v[e + half];
                const float keep = hb ? v[e + half] : v[e];
                v[e] = keep + __shfl_xor_sync(0xffffffffu, send, lanebit);
            }
        }
        n = half;
    }
    return v[0];
}
template <unsigned GLMASK> __device__ __forceinline__ float group_max(float v) {
#pragma unroll
    for (int b = 4; b >= 0; --b)
        if (GLMASK >> b & 1u) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1 << b));
    return v;
}

// PADDED: the caller's dimension Dr is smaller than the engine's D (rows Dr..D-1 are an identity block); the exact-size
// instantiation carries no run-time dimension at all.
template <int D, int BS, int WARPS, int MINB, bool TMA, int PF, bool PADDED>
__global__ void __launch_bounds__(WARPS * 32, MINB) local_step_fast_kernel(const FastParams p) {
    using G = FastGeom<D, BS>;
    constexpr int ROWS = G::ROWS, GPW = 32 / BS, PPC = WARPS * GPW, LD = G::LD, REC = G::REC, TS = G::TS;
    constexpr int GS = G::GS;
    constexpr unsigned FULL = 0xffffffffu;
    static_assert(D % 16 == 0 && BS * ROWS == D && BS % 4 == 0 && 32 % BS == 0, "unsupported (D, BS)");

    extern __shared__ __align__(128) unsigned char smraw[];
    float* stage = reinterpret_cast<float*>(smraw);              // [2][REC]
    float* gsm = stage + 2 * REC;                                // [PPC][GS]
    float* ksm = gsm + PPC * GS;                                 // [PPC][K][3]
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ double cta_acc[4];

    constexpr unsigned GLMASK = G::GLMASK;
    static_assert(bits_extract(GLMASK, GLMASK) == BS - 1, "GLMASK must have log2(BS) bits");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gl = (int)bits_extract((unsigned)lane, GLMASK);
    const int grp = warp * GPW + (int)bits_extract((unsigned)lane, ~GLMASK & 31u);
    // value of group-lane l of this lane's pair: ONE shfl.idx whose segment mask (c[12:8]) keeps the pair bits of the lane id
    // and takes the gl bits from the immediate source — PTX allows any 5-bit mask there, not only the 2^n - 1 "width" forms
    auto group_bcast = [&](float v, int l) {
        float out;
        asm volatile("shfl.sync.idx.b32 %0, %1, %2, %3, 0xffffffff;"
                     : "=f"(out)
                     : "f"(v), "r"((int)bits_deposit((unsigned)l, GLMASK)), "r"((int)(((~GLMASK & 31u) << 8) | 0x1fu)));
        return out;
    };
    float* col = gsm + grp * GS;          // [2][D]
    float* vec = col + 2 * D;             // [D]
    float* ybf = vec + D;                 // [BS]
    float* tbf = ybf + (BS < 4 ? 4 : BS); // [BS][TS]
    float* mub = tbf + BS * TS;           // [D] mu1 of this group's point
    float* p1b = mub + D;                 // [D] p1
    float* xb = p1b + D;                  // [D] sample of the running Gumbel arg-max
    float* ab = xb + D;                   // [D] a = L^-1 P2 d of the current pair
    float* ib = ab + D;                   // [D] 1 / L_jj of the current pair
    float* kst = ksm + (size_t)grp * p.K * 3;
    const int K = p.K, S = p.S;
    const int Dr = PADDED ? p.Dr : D;

    if (tid == 0) {
        cta_acc[0] = cta_acc[1] = cta_acc[2] = cta_acc[3] = 0.0;
        if (TMA) {
            mbar_init(&bars[0], 1);
            mbar_init(&bars[1], 1);
            fence_barrier_init();
        }
    }
    __syncthreads();
    uint32_t phase0 = 0, phase1 = 0;

    auto issue_load = [&](int k, int buf) {
        // called by every thread; TMA: one elected thread issues a bulk copy, else: cooperative cp.async
        const float* src = p.recs + (size_t)k * REC;
        float* dst = stage + buf * REC;
        if (TMA) {
            if (tid == 0) {
                mbar_expect_tx(&bars[buf], REC * 4);
                bulk_g2s(dst, src, REC * 4, &bars[buf]);
            }
        } else {
            for (int e = tid; e < REC / 4; e += WARPS * 32) {
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + 4 * e)), "l"(src + 4 * e)
                             : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    };

    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int64_t n_raw = tile * PPC + grp;
        const bool active = n_raw < p.N;
        const int64_t n = active ? n_raw : p.N - 1;

#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const int row = r * BS + gl;
            // padded rows (row >= Dr): P~ gets an identity block there (the packed P2 carries the 1), mu1 = 0
            const bool real = row < Dr;
            const float p1v = real ? -2.f * p.eta2d[n * Dr + row] : 0.f;
            p1b[row] = p1v;
            mub[row] = real ? p.eta1[n * Dr + row] / p1v : 0.f;
            xb[row] = 0.f;
        }
        float best = -CUDART_INF_F;
        int zbest = 0;
        int bad = 0;

        issue_load(0, 0);
        for (int k = 0; k < K; ++k) {
            const int buf = k & 1;
            // stage[buf] holds record k once its copy has landed; the barrier also guarantees that every warp is
            // done with stage[buf^1] (read at k-1) before record k+1 is copied over it
            if (TMA) {
                mbar_wait(&bars[buf], buf ? phase1 : phase0);
                if (buf) phase1 ^= 1; else phase0 ^= 1;
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            __syncthreads();
            if (k + 1 < K) issue_load(k + 1, buf ^ 1);

            const float* rec = stage + buf * REC;
            const float* P2s = rec;
            const float* Ws = rec + D * LD;
            const float* mu2 = rec + 2 * D * LD;
            const float* mth = mu2 + D;
            const float* scl = mth + D;

            // ---------------- phase 1: d, g1 = P1 d, g = P2 d, A <- P2 rows
            float g[ROWS], g1[ROWS];
            float A[ROWS][D];
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                const int row = r * BS + gl;
                const float d = mub[row] - mu2[row];
                g1[r] = p1b[row] * d;
                vec[row] = d;
            }
            __syncwarp();
            static_for<0, ROWS>([&](auto rc) {
                constexpr int r = decltype(rc)::value;
                const float4* prow = reinterpret_cast<const float4*>(P2s + (r * BS + gl) * LD);
                const float4* dv4 = reinterpret_cast<const float4*>(vec);
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                for (int c4 = 0; c4 < D / 4; ++c4) {
                    const float4 pv = prow[c4];
                    const float4 dv = dv4[c4];
                    ffma2(s0, s1, pv.x, pv.y, dv.x, dv.y);
                    ffma2(s2, s3, pv.z, pv.w, dv.z, dv.w);
                    if (4 * c4 < (r + 1) * BS) {
                        A[r][4 * c4 + 0] = pv.x;
                        A[r][4 * c4 + 1] = pv.y;
                        A[r][4 * c4 + 2] = pv.z;
                        A[r][4 * c4 + 3] = pv.w;
                    }
                }
                g[r] = (s0 + s1) + (s2 + s3);
                // P~ = P2 + diag(p1): the diagonal entry of this lane's row sits in column r*BS + gl
                const float p1r = p1b[r * BS + gl];
#pragma unroll
                for (int cc = 0; cc < BS; ++cc) A[r][r * BS + cc] += (gl == cc) ? p1r : 0.f;
            });

            // ---------------- phase 2: right-looking factorisation with the two forward substitutions riding along.
            // Square-root-free form P~ = Lu diag(d) Lu^T (Lu unit lower, L = Lu diag(sqrt d)): the column is published RAW and the
            // pivot-row values of the right-hand sides are broadcast RAW, so neither waits for the MUFU chain of the pivot —
            // per column only the multipliers m = A[.][j] / d_j depend on it (SHFL -> RCP -> Newton -> FMUL), and that chain runs
            // beside the store -> load round trip of the column instead of in front of it.  a = L^-1 g = z / sqrt d is formed per
            // column for the samples; a . a1 = sum z z1 / d.
            constexpr bool KEEP_Z = BS == 4;
            float q = 0.f, hl2 = 0.f;                                     // a . a1 and sum_j log2(d_j)
            float iq[4], yq[4];
            static_for<0, D>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                constexpr int rj = j / BS, lj = j % BS;
                const float piv = group_bcast(A[rj][j], lj);
                const float zj = group_bcast(g[rj], lj);
                const float z1j = group_bcast(g1[rj], lj);
                float* cw = col + (j & 1) * D;
#pragma unroll
                for (int r = rj; r < ROWS; ++r) cw[r * BS + gl] = A[r][j];
                float rcp = rcp_approx(piv);                             // bare MUFU.RCP (pivots are O(1): no denormals)
                rcp = fmaf(rcp, fmaf(-piv, rcp, 1.f), rcp);              // one Newton step: ~0.5 ulp
                // d_j (and z_j, see KEEP_Z) go to shared memory for the per-row epilogue below (group-uniform: four columns are
                // batched into one 128-bit store by lane 0).  1 / sqrt(d_j), a_j and log d_j are NOT formed per column by all BS
                // lanes redundantly: every lane does it for its own ROWS rows after the loop.
                // KEEP_Z (BS = 4): z, z1 also stay in g, g1 of the row's owner (its own pivot-row update is masked out), so a . a1
                // is per-row work too; at BS = 16 that keeps 8 more registers live through the loop and costs more in spills
                // than it saves (measured 0.329 against 0.336), so z_j is staged and a . a1 accumulated per column there.
                iq[j & 3] = piv;
                if constexpr (!KEEP_Z) {
                    q = fmaf(zj * rcp, z1j, q);
                    yq[j & 3] = zj;
                }
                if constexpr ((j & 3) == 3) {
                    if (gl == 0) {
                        *reinterpret_cast<float4*>(ib + j - 3) = make_float4(iq[0], iq[1], iq[2], iq[3]);
                        if constexpr (!KEEP_Z) *reinterpret_cast<float4*>(ab + j - 3) = make_float4(yq[0], yq[1], yq[2], yq[3]);
                    }
                }
#pragma unroll
                for (int r = rj; r < ROWS; ++r) {
                    float m = A[r][j] * rcp;                             // Lu[r][j] (the owner's diagonal entry becomes 1)
                    if (KEEP_Z && r == rj) m = (gl > lj) ? m : 0.f;      // rows <= j of this block row are finished: keep their z
                    A[r][j] = m;
                    ffma2_bcast(g[r], g1[r], -m, zj, z1j);
                }
                __syncwarp();
                // trailing update A[r][c] -= Lu[r][j] * A[c][j] (raw column), c > j, as packed pairs (c0, c0+1)
                // the 128-bit broadcast reads are software-pipelined PF chunks ahead of their FFMA2s (ptxas otherwise
                // recycles one 4-register buffer and exposes the shared-memory latency on every chunk)
                constexpr int C4B = (j + 1) / 4, C4E = D / 4;
                const float4* cr4 = reinterpret_cast<const float4*>(cw);
                float4 pf[PF];
#pragma unroll
                for (int t = 0; t < PF; ++t)
                    if (C4B + t < C4E) pf[t] = cr4[C4B + t];
                static_for<C4B, C4E>([&](auto c4c) {
                    constexpr int c4 = decltype(c4c)::value;
                    const float4 cv = pf[(c4 - C4B) % PF];
                    if constexpr (c4 + PF < C4E) pf[(c4 - C4B) % PF] = cr4[c4 + PF];
                    const float cvv[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int c0 = 4 * c4 + 2 * h;
                        if (c0 > j) {
#pragma unroll
                            for (int r = rj; r < ROWS; ++r)
                                if (c0 < (r + 1) * BS)
                                    ffma2_bcast(A[r][c0], A[r][c0 + 1], -A[r][j], cvv[2 * h], cvv[2 * h + 1]);
                        } else if (c0 + 1 > j) {
#pragma unroll
                            for (int r = rj; r < ROWS; ++r)
                                if (c0 + 1 < (r + 1) * BS) A[r][c0 + 1] = fmaf(-A[r][j], cvv[2 * h + 1], A[r][c0 + 1]);
                        }
                    }
                });
            });
            // own rows: 1 / L_ii = 1 / sqrt(d_i) (one Newton step: ~0.5 ulp), a_i = z_i / sqrt(d_i), a . a1, sum of log2 d_i
            float ar[ROWS], rsr[ROWS];
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                const float d = ib[r * BS + gl];
                float rs = rsqrt_approx(d);                              // bare MUFU.RSQ (pivots are O(1): no denormals)
                rs = rs * fmaf(-0.5f * d * rs, rs, 1.5f);
                rsr[r] = rs;
                if constexpr (KEEP_Z) {
                    ar[r] = g[r] * rs;
                    q = fmaf(ar[r], g1[r] * rs, q);
                } else {
                    ar[r] = ab[r * BS + gl] * rs;
                }
                hl2 += lg2_approx(d);                                    // bare MUFU.LG2 (rel. error 2^-22; NaN flags a bad pivot)
            }
            if constexpr (KEEP_Z) q = group_sum<GLMASK>(q);
            hl2 = group_sum<GLMASK>(hl2);
            const float hld = 0.5f * (float)VMP_LOG_2 * hl2;           // sum_i log L_ii
            bad |= !(fabsf(hld) < CUDART_INF_F);                        // a non-positive pivot shows up as NaN / inf
            const float score = scl[0] - 0.5f * q + 0.5f * scl[1] - hld;

            // ---------------- phase 3: samples, ELBO terms
            const uint64_t pair = (uint64_t)n * K + k;                       // index into caller buffers
            const uint64_t gpair = pair + p.pair_offset;                      // key of the in-kernel noise stream
            float snum = 0.f, sden = 0.f;
            float x0[ROWS];
            float u_pair = 0.f;                                           // Gumbel uniform of the pair (spare bits of its first Philox block)
            for (int s = 0; s < S; ++s) {
                float u_spare = 0.f;
                float w[ROWS], y[ROWS];
                if (p.noise != nullptr) {
#pragma unroll
                    for (int r = 0; r < ROWS; ++r)
                        w[r] = (r * BS + gl < Dr) ? p.noise[(pair * Dr + r * BS + gl) * (uint64_t)S + s] : 0.f;
                } else {
                    __syncwarp();
                    for (int qd = gl; qd < D / 4; qd += BS) {
                        float sp;
                        reinterpret_cast<float4*>(vec)[qd] = philox_normal4(p.seed, gpair, (uint32_t)s, (uint32_t)qd, &sp);
                        if (qd < BS) u_spare = sp;
                    }
                    if (s == 0) u_pair = group_bcast(u_spare, 0);        // block (s = 0, q = 0) is group-lane 0's first call
                    __syncwarp();
#pragma unroll
                    for (int r = 0; r < ROWS; ++r) w[r] = (r * BS + gl < Dr) ? vec[r * BS + gl] : 0.f;
                }
                float e2 = 0.f;
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    e2 = fmaf(w[r], w[r], e2);
                    w[r] = (w[r] - ar[r]) * rsr[r];                     // L^T y = w  <=>  Lu^T y = w / sqrt d  (unit upper)
                    y[r] = 0.f;
                }
                // back substitution L^T y = w, block rows from the bottom
                if constexpr (BS == 4) {
                    // Row-owner form without any transposition: the owner of row R, once y_R is known, adds L[R][c] y_R to its
                    // PRIVATE partial sums pw[c], c < R; a block of four columns is completed by a reduce-scatter over the
                    // four lanes of the pair (3 shuffles) when the sweep reaches it, the in-block terms by 5 more.  No shared
                    // memory: the 36 (D=32) / 10 (D=16) 4x4 block transposes of the tile form were ~22 % of this engine's
                    // shared-memory wavefronts.
                    constexpr int HI = 31 - __builtin_clz(GLMASK), LO = __builtin_ctz(GLMASK);   // lane bits of gl bit 1 / bit 0
                    const bool ghi = (gl & 2) != 0, glo = (gl & 1) != 0;
                    float pw[D - BS];
#pragma unroll
                    for (int c = 0; c < D - BS; ++c) pw[c] = 0.f;
                    static_for_down<ROWS>([&](auto rbc) {
                        constexpr int rb = decltype(rbc)::value;
                        float wv = w[rb];
                        if constexpr (rb < ROWS - 1) {
                            // lane gl <- sum over the four lanes of pw[rb*4 + gl]
                            const float s0 = ghi ? pw[rb * 4 + 0] : pw[rb * 4 + 2], s1 = ghi ? pw[rb * 4 + 1] : pw[rb * 4 + 3];
                            float k0 = ghi ? pw[rb * 4 + 2] : pw[rb * 4 + 0], k1 = ghi ? pw[rb * 4 + 3] : pw[rb * 4 + 1];
                            k0 += __shfl_xor_sync(FULL, s0, 1 << HI);
                            k1 += __shfl_xor_sync(FULL, s1, 1 << HI);
                            const float s2 = glo ? k0 : k1;
                            float kk = glo ? k1 : k0;
                            kk += __shfl_xor_sync(FULL, s2, 1 << LO);
                            wv -= kk;
                        }
                        // rows 3..0 of the block: lane i finishes y_i, its in-block terms L[i][c] y_i wait in pd[c], c < i
                        float pd0 = 0.f, pd1 = 0.f, pd2 = 0.f;
                        {   // i = 3
                            const float ym = (gl == 3) ? wv : 0.f;
                            y[rb] = (gl == 3) ? ym : y[rb];
                            pd2 = A[rb][rb * 4 + 2] * ym;
                            pd1 = A[rb][rb * 4 + 1] * ym;
                            pd0 = A[rb][rb * 4 + 0] * ym;
                        }
                        {   // i = 2: lane 2 needs pd2 of lane 3
                            const float t = group_bcast(pd2, 3);
                            const float ym = (gl == 2) ? (wv - t) : 0.f;
                            y[rb] = (gl == 2) ? ym : y[rb];
                            pd1 = fmaf(A[rb][rb * 4 + 1], ym, pd1);
                            pd0 = fmaf(A[rb][rb * 4 + 0], ym, pd0);
                        }
                        {   // i = 1: lane 1 needs pd1 of lanes 2 and 3 (lanes 0, 1 hold 0)
                            float t = pd1 + __shfl_xor_sync(FULL, pd1, 1 << HI);
                            t += __shfl_xor_sync(FULL, t, 1 << LO);
                            const float ym = (gl == 1) ? (wv - t) : 0.f;
                            y[rb] = (gl == 1) ? ym : y[rb];
                            pd0 = fmaf(A[rb][rb * 4 + 0], ym, pd0);
                        }
                        {   // i = 0
                            float t = pd0 + __shfl_xor_sync(FULL, pd0, 1 << HI);
                            t += __shfl_xor_sync(FULL, t, 1 << LO);
                            if (gl == 0) y[rb] = wv - t;
                        }
                        // every lane now owns its y of this block row: its terms for all earlier columns
                        if constexpr (rb > 0) {
#pragma unroll
                            for (int c = 0; c < rb * 4; c += 2) ffma2_bcast(pw[c], pw[c + 1], y[rb], A[rb][c], A[rb][c + 1]);
                        }
                    });
                } else {
                    static_for_down<ROWS>([&](auto rbc) {
                        constexpr int rb = decltype(rbc)::value;
                        if constexpr (rb < ROWS - 1) {
                            // terms of the rows below for this block's columns: every lane multiplies its own rows (private
                            // sums, lazy: only BS of them are live), then a reduce-scatter over the pair's lanes (BS - 1
                            // shuffles) hands column rb*BS + gl to lane gl — instead of ROWS-1-rb tile transposes
                            float pv[BS];
#pragma unroll
                            for (int e = 0; e < BS; ++e) pv[e] = 0.f;
                            static_for<rb + 1, ROWS>([&](auto rsc) {
                                constexpr int rs = decltype(rsc)::value;
#pragma unroll
                                for (int e = 0; e < BS; e += 2)
                                    ffma2_bcast(pv[e], pv[e + 1], y[rs], A[rs][rb * BS + e], A[rs][rb * BS + e + 1]);
                            });
                            w[rb] -= group_reduce_scatter<BS, GLMASK>(pv, gl);
                        }
                        __syncwarp();
    #pragma unroll
                        for (int q4 = 0; q4 < BS / 4; ++q4)
                            reinterpret_cast<float4*>(tbf + gl * TS)[q4] =
                                make_float4(A[rb][rb * BS + 4 * q4], A[rb][rb * BS + 4 * q4 + 1], A[rb][rb * BS + 4 * q4 + 2],
                                            A[rb][rb * BS + 4 * q4 + 3]);
                        __syncwarp();
    #pragma unroll
                        for (int i = BS - 1; i >= 0; --i) {
                            const float t = tbf[i * TS + gl];               // Lu[rb*BS+i][rb*BS+gl]
                            const float yi = group_bcast(w[rb], i);
                            y[rb] = (gl == i) ? yi : y[rb];
                            w[rb] = fmaf(-t, yi, w[rb]);
                        }
                    });
                }
                // x = mu1 + y ; publish x - m_theta for the quadratic form
                float x[ROWS];
                __syncwarp();
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    x[r] = mub[r * BS + gl] + y[r];
                    vec[r * BS + gl] = x[r] - mth[r * BS + gl];
                    if (s == 0) x0[r] = x[r];
                }
                if (p.x_k_samples != nullptr && active) {
#pragma unroll
                    for (int r = 0; r < ROWS; ++r)
                        if (r * BS + gl < Dr) p.x_k_samples[((pair * S) + s) * (uint64_t)Dr + r * BS + gl] = x[r];
                }
                __syncwarp();
                float q2 = 0.f;
                static_for<0, ROWS>([&](auto rc) {
                    constexpr int r = decltype(rc)::value;
                    const float4* wrow = reinterpret_cast<const float4*>(Ws + (r * BS + gl) * LD);
                    const float4* xv4 = reinterpret_cast<const float4*>(vec);
                    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
#pragma unroll
                    for (int c4 = 0; c4 < (r + 1) * BS / 4; ++c4) {
                        const float4 wv = wrow[c4];
                        const float4 xv = xv4[c4];
                        ffma2(t0, t1, wv.x, wv.y, xv.x, xv.y);
                        ffma2(t2, t3, wv.z, wv.w, xv.z, xv.w);
                    }
                    const float t = (t0 + t1) + (t2 + t3);
                    q2 = fmaf(t, t, q2);
                });
                e2 = group_sum<GLMASK>(e2);
                q2 = group_sum<GLMASK>(q2);
                snum += -0.5f * e2;
                sden += (p.den_mode == VMP_DEN_GAUSS) ? (scl[2] - 0.5f * q2)
                                                      : (scl[2] - 0.5f * (scl[3] + (float)Dr) * log1pf(q2 / scl[3]));
            }
            // ---------------- per-pair results
            if (gl == 0) {
                kst[3 * k + 0] = score;
                kst[3 * k + 1] = snum / (float)S + hld - 0.5f * (float)VMP_LOG_2PI * (float)Dr;
                kst[3 * k + 2] = sden / (float)S;
            }
            // online Gumbel-max draw of z_n (tf.multinomial GPU algorithm): keep the running arg-max and its sample
            const float u = p.gum_u != nullptr ? p.gum_u[pair] : (p.noise != nullptr ? philox_uniform_pair(p.seed, gpair) : u_pair);
            const float cand = score + gumbel_from_uniform<float>(u);
            if (cand > best) {
                best = cand;
                zbest = k;
#pragma unroll
                for (int r = 0; r < ROWS; ++r) xb[r * BS + gl] = x0[r];
            }
        }

        // ---------------- per-point epilogue: log-sum-exp over K, log_r, ELBO partials, selected sample
        __syncwarp();
        float mx = -CUDART_INF_F;
        for (int k = gl; k < K; k += BS) mx = fmaxf(mx, kst[3 * k]);
        mx = group_max<GLMASK>(mx);
        double se = 0.0;
        for (int k = gl; k < K; k += BS) se += (double)expf(kst[3 * k] - mx);
        se = group_sum_d<GLMASK>(se);
        const float lse = mx + (float)log(se);
        double en = 0.0, ed = 0.0;
        if (active) {
            for (int k = gl; k < K; k += BS) {
                const float lr = kst[3 * k] - lse;
                p.log_r[n * K + k] = lr;
                const double r = (double)expf(lr);
                en += r * ((double)kst[3 * k + 1] + (double)lr);
                ed += r * (double)kst[3 * k + 2];
            }
            if (p.x_sample != nullptr) {
#pragma unroll
                for (int r = 0; r < ROWS; ++r)
                    if (r * BS + gl < Dr) p.x_sample[n * Dr + r * BS + gl] = xb[r * BS + gl];
            }
            if (p.z != nullptr && gl == 0) p.z[n] = zbest;
        }
        en = group_sum_d<GLMASK>(en);
        ed = group_sum_d<GLMASK>(ed);
        if (gl == 0 && active) {
            atomicAdd(&cta_acc[0], en);
            atomicAdd(&cta_acc[1], ed);
            if (bad) atomicAdd(&cta_acc[3], 1.0);
        }
        __syncthreads();   // stage buffers and kst are reused by the next tile
    }
    __syncthreads();
    if (tid == 0) {
        atomicAdd(p.elbo_acc + 0, cta_acc[0]);
        atomicAdd(p.elbo_acc + 1, cta_acc[1]);
        atomicAdd(p.elbo_acc + 2, cta_acc[0] - cta_acc[1]);
        if (cta_acc[3] != 0.0) atomicAdd(p.elbo_acc + 3, cta_acc[3]);
    }
}

template <int D, int BS, bool TMA, int PF, bool PADDED>
static int launch_fast_t(const FastParams& p0, cudaStream_t st) {
    using G = FastGeom<D, BS>;
    constexpr int WARPS = FastLaunch<D, BS>::WARPS, MINB = FastLaunch<D, BS>::MINB;
    constexpr int PPC = WARPS * (32 / G::BS);
    FastParams p = p0;
    p.ntiles = (p.N + PPC - 1) / PPC;
    const size_t smem = fast_smem_bytes<D, BS>(p.K);
    auto kern = local_step_fast_kernel<D, BS, WARPS, MINB, TMA, PF, PADDED>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 148, occ = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem);
    if (occ < 1) occ = 1;
    int64_t grid = (int64_t)sms * occ;
    if (grid > p.ntiles) grid = p.ntiles;
    kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(p);
    return launch_status();
}

#define VMP_FAST_INSTANTIATE(DD, BB)                                                  \
    template <> int launch_fast<DD, BB>(const FastParams& p, cudaStream_t st) {      \
        return p.Dr == DD ? launch_fast_t<DD, BB, true, 4, false>(p, st)             \
                          : launch_fast_t<DD, BB, true, 4, true>(p, st);             \
    }
#endif  // VMP_FAST_IMPL

}  // namespace vmp
