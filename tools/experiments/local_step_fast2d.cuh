// local_step_fast2d.cuh — 2-D cyclic variant of the register-resident group engine (fp32, D = 64).
//
// Why: in the row-owned engine (local_step_fast.cuh) every lane reads the WHOLE pivot column each step but uses a
// value for only ~1.3 of its 4 rows on average (the triangle), and the shared-memory wavefront pipe is the limiter
// (ncu: 2215 wavefronts per warp and component, 46 % of them these column broadcasts).  Here the 16 lanes of a pair
// form a 4 x 4 grid: lane (pr, pc) owns A(i, c) with i = 4a + pr, c = 4b + pc as register A[a][b], b <= a.  A step
// then needs the pivot-column entries of the lane's ROW class and of its COLUMN class only — half the values — and a
// half-warp reads four distinct 16-byte chunks in one wavefront.  The rest follows tools/engine2d_model.py (validated
// lane by lane in numpy):
//   * deferred scaling: step j publishes the UNSCALED column (zero for finished rows) BEFORE the pivot is known, so
//     the publish -> read latency overlaps the rsqrt; A[a][b] -= (s_a inv^2) s_b; columns are scaled by 1/L_jj when
//     their block of four is complete;
//   * the two right-hand sides ride along as an extra row a = NB (lanes pr = 0: g = P2 d, pr = 1: g1 = P1 d), so the
//     forward substitutions are part of the factorisation;
//   * g = P2 d is a symmetric mat-vec on the lower-triangular registers (column sums reduced over pr, row sums over pc);
//   * back substitution by 4-column blocks with per-lane partial sums and shuffles only (no shared-memory transposes);
//   * |W (x - m)|^2 as a row-class partial mat-vec reduced over pc.
// The staged per-component record is lane-major (each lane's registers are contiguous 16-byte chunks):
//   P2L[NCH][16][4] | WL[NCH][16][4] | mu2 by residue [4][RS] | m by residue [4][RS] | 8 scalars.
#pragma once
#include "local_step_fast.cuh"

namespace vmp {

template <int D> struct Fast2dGeom {
    static constexpr int NB = D / 4;                 // 4x4 blocks per dimension
    static constexpr int RS = NB + 4;                // row stride of a residue-class vector ([4][RS]: r -> r*RS + t)
    static constexpr int NCH = 4 * ((NB / 4) * (NB / 4 + 1) / 2);   // 16-byte chunks per lane, rows padded to 4 entries
    static constexpr int MAT = NCH * 16 * 4;
    static constexpr int REC = 2 * MAT + 2 * 4 * RS + 8;
    // group scratch: cb[2][4*RS] | mures | p1res | dres | g1res | epsres | xbres (each 4*RS) | ib[D]
    static constexpr int GS_RAW = 8 * 4 * RS + D;
    static constexpr int GS = ((GS_RAW - 16 + 31) / 32) * 32 + 16;
    __host__ __device__ static constexpr int chbase(int a) {      // first chunk of row a
        const int g = a / 4;
        return 4 * (g * (g + 1) / 2) + (a - 4 * g) * (g + 1);
    }
};

inline size_t fast2d_smem_bytes(int K) {
    using G = Fast2dGeom<64>;
    return sizeof(float) * (2 * (size_t)G::REC + 16 * (size_t)G::GS + 16 * (size_t)K * 3);
}
inline int fast2d_rec_len() { return Fast2dGeom<64>::REC; }

int launch_fast2d_64(const FastParams& p, cudaStream_t st);
void launch_pack_fast2d_records(int K, const float* phi_rec, const float* theta_rec, float* out, cudaStream_t st);

#ifdef VMP_FAST2D_IMPL
// (d0, d1) = (a0, a1) * s  (packed FP32x2 multiply)
__device__ __forceinline__ void fmul2_bcast(float& d0, float& d1, float a0, float a1, float s) {
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2,%3};\n\tmov.b64 rb, {%4,%4};\n\t"
        "mul.rn.f32x2 rc, ra, rb;\n\tmov.b64 {%0,%1}, rc;\n\t}"
        : "=f"(d0), "=f"(d1)
        : "f"(a0), "f"(a1), "f"(s));
}

template <int D, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) local_step_fast2d_kernel(const FastParams p) {
    using G = Fast2dGeom<D>;
    constexpr int NB = G::NB, RS = G::RS, MAT = G::MAT, REC = G::REC, GS = G::GS, PPC = WARPS * 2;
    constexpr unsigned FULL = 0xffffffffu;
    static_assert(NB % 4 == 0, "D must be a multiple of 16");

    extern __shared__ __align__(128) unsigned char smraw[];
    float* stage = reinterpret_cast<float*>(smraw);              // [2][REC]
    float* gsm = stage + 2 * REC;                                // [PPC][GS]
    float* ksm = gsm + PPC * GS;                                 // [PPC][K][3]
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ double cta_acc[4];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gl = lane & 15, pr = gl >> 2, pc = gl & 3;
    const int grp = warp * 2 + (lane >> 4);
    float* cb = gsm + grp * GS;           // [2][4*RS] published pivot column, by row residue
    float* mures = cb + 2 * 4 * RS;       // mu1 of this group's point, by residue
    float* p1res = mures + 4 * RS;
    float* dres = p1res + 4 * RS;
    float* g1res = dres + 4 * RS;
    float* epsres = g1res + 4 * RS;
    float* xbres = epsres + 4 * RS;       // sample of the running Gumbel arg-max
    float* ib = xbres + 4 * RS;           // [D] 1 / L_jj of the current pair
    float* kst = ksm + (size_t)grp * p.K * 3;
    const int K = p.K, S = p.S;
    const float diagf = (pr == pc) ? 1.f : 0.f;

    if (tid == 0) {
        cta_acc[0] = cta_acc[1] = cta_acc[2] = cta_acc[3] = 0.0;
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_barrier_init();
    }
    __syncthreads();
    uint32_t phase0 = 0, phase1 = 0;
    auto issue_load = [&](int k, int buf) {
        if (tid == 0) {
            mbar_expect_tx(&bars[buf], REC * 4);
            bulk_g2s(stage + buf * REC, p.recs + (size_t)k * REC, REC * 4, &bars[buf]);
        }
    };
    // this lane's slice of a 64-vector in residue layout: residue r = gl >> 2, entries t = 4 (gl & 3) .. +3
    const int vr = gl >> 2, vt = 4 * (gl & 3);

    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int64_t n_raw = tile * PPC + grp;
        const bool active = n_raw < p.N;
        const int64_t n = active ? n_raw : p.N - 1;
        {
            float4 mu4, p14;
            float* mu = reinterpret_cast<float*>(&mu4);
            float* pp = reinterpret_cast<float*>(&p14);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int e = 4 * (vt + i) + vr;
                const float p1v = -2.f * p.eta2d[n * D + e];
                pp[i] = p1v;
                mu[i] = p.eta1[n * D + e] / p1v;
            }
            *reinterpret_cast<float4*>(mures + vr * RS + vt) = mu4;
            *reinterpret_cast<float4*>(p1res + vr * RS + vt) = p14;
            *reinterpret_cast<float4*>(xbres + vr * RS + vt) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float best = -CUDART_INF_F;
        int zbest = 0;
        int bad = 0;

        issue_load(0, 0);
        for (int k = 0; k < K; ++k) {
            const int buf = k & 1;
            mbar_wait(&bars[buf], buf ? phase1 : phase0);
            if (buf) phase1 ^= 1; else phase0 ^= 1;
            __syncthreads();
            if (k + 1 < K) issue_load(k + 1, buf ^ 1);

            const float* rec = stage + buf * REC;
            const float4* P2L = reinterpret_cast<const float4*>(rec) + gl;            // chunk ch at P2L[ch * 16]
            const float4* WL = reinterpret_cast<const float4*>(rec + MAT) + gl;
            const float* mu2res = rec + 2 * MAT;
            const float* mthres = mu2res + 4 * RS;
            const float* scl = mthres + 4 * RS;

            // ---------------- phase 1: d, g1 = P1 d by residue; A <- P2 (lower), g = P2 d as a symmetric mat-vec
            {
                const float4 mu4 = *reinterpret_cast<const float4*>(mures + vr * RS + vt);
                const float4 m24 = *reinterpret_cast<const float4*>(mu2res + vr * RS + vt);
                const float4 p14 = *reinterpret_cast<const float4*>(p1res + vr * RS + vt);
                const float4 d4 = make_float4(mu4.x - m24.x, mu4.y - m24.y, mu4.z - m24.z, mu4.w - m24.w);
                *reinterpret_cast<float4*>(dres + vr * RS + vt) = d4;
                *reinterpret_cast<float4*>(g1res + vr * RS + vt) = make_float4(p14.x * d4.x, p14.y * d4.y, p14.z * d4.z, p14.w * d4.w);
            }
            __syncwarp();
            float A[NB + 1][NB];
            {
                float dr[NB], dc[NB], colsum[NB], rowsum[NB];      // dr / dc are read block by block (short live ranges)
#pragma unroll
                for (int b = 0; b < NB; ++b) { colsum[b] = 0.f; rowsum[b] = 0.f; }
                static_for<0, NB>([&](auto ac) {
                    constexpr int a = decltype(ac)::value;
                    if constexpr (a % 4 == 0) {
                        const float4 a4 = *reinterpret_cast<const float4*>(dres + pr * RS + a);
                        const float4 b4 = *reinterpret_cast<const float4*>(dres + pc * RS + a);
                        dr[a] = a4.x; dr[a + 1] = a4.y; dr[a + 2] = a4.z; dr[a + 3] = a4.w;
                        dc[a] = b4.x; dc[a + 1] = b4.y; dc[a + 2] = b4.z; dc[a + 3] = b4.w;
                    }
                    static_for<0, a / 4 + 1>([&](auto tc) {
                        constexpr int t = decltype(tc)::value;
                        const float4 e4 = P2L[(G::chbase(a) + t) * 16];
                        const float ev[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int b0 = 4 * t + 2 * h;
                            if (b0 + 1 <= a) {
                                A[a][b0] = ev[2 * h];
                                A[a][b0 + 1] = ev[2 * h + 1];
                                ffma2_bcast(colsum[b0], colsum[b0 + 1], dr[a], ev[2 * h], ev[2 * h + 1]);
                                rowsum[a] = fmaf(ev[2 * h], dc[b0], rowsum[a]);
                                rowsum[a] = fmaf(ev[2 * h + 1], (b0 + 1 == a) ? dc[b0 + 1] * (1.f - diagf) : dc[b0 + 1], rowsum[a]);
                            } else if (b0 <= a) {
                                A[a][b0] = ev[2 * h];
                                colsum[b0] = fmaf(ev[2 * h], dr[a], colsum[b0]);
                                rowsum[a] = fmaf(ev[2 * h], (b0 == a) ? dc[b0] * (1.f - diagf) : dc[b0], rowsum[a]);
                            }
                        }
                    });
                });
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    colsum[b] += __shfl_xor_sync(FULL, colsum[b], 4, 16);
                    colsum[b] += __shfl_xor_sync(FULL, colsum[b], 8, 16);
                    rowsum[b] += __shfl_xor_sync(FULL, rowsum[b], 1, 16);
                    rowsum[b] += __shfl_xor_sync(FULL, rowsum[b], 2, 16);
                }
#pragma unroll
                for (int t = 0; t < NB / 4; ++t) {
                    const float4 g4 = *reinterpret_cast<const float4*>(g1res + pc * RS + 4 * t);
                    const float4 p4 = *reinterpret_cast<const float4*>(p1res + pr * RS + 4 * t);
                    const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
                    const float pv[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int b = 4 * t + i;
                        const float rowg = __shfl_sync(FULL, rowsum[b], pc * 4, 16);      // row 4b+pc lives on lanes pr' = pc
                        A[NB][b] = (pr == 0) ? colsum[b] + rowg : ((pr == 1) ? gv[i] : 0.f);
                        A[b][b] = fmaf(diagf, pv[i], A[b][b]);                             // P~ = P2 + diag(p1)
                    }
                }
            }

            // ---------------- phase 2: right-looking Cholesky, deferred scaling, RHS row riding along
            float q = 0.f, hl2 = 0.f, myinv = 0.f;
            float invq[4];
            static_for<0, D>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                constexpr int ja = j / 4, jr = j % 4, t0 = ja / 4;
                float* cbw = cb + (j & 1) * 4 * RS;
                if (pc == jr) {
                    static_for<t0, NB / 4 + 1>([&](auto tc) {
                        constexpr int t = decltype(tc)::value;
                        float v[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int a = 4 * t + i;
                            if (a < ja || a > NB) v[i] = 0.f;
                            else if (a == ja) v[i] = (pr > jr) ? A[ja][ja] : 0.f;
                            else v[i] = A[a][ja];
                        }
                        *reinterpret_cast<float4*>(cbw + pr * RS + 4 * t) = make_float4(v[0], v[1], v[2], v[3]);
                    });
                }
                const float piv = __shfl_sync(FULL, A[ja][ja], 5 * jr, 16);
                float inv = rsqrt_approx(piv);
                inv = inv * fmaf(-0.5f * piv * inv, inv, 1.5f);
                hl2 += lg2_approx(piv);
                const float gj = __shfl_sync(FULL, A[NB][ja], jr, 16);
                const float g1j = __shfl_sync(FULL, A[NB][ja], 4 + jr, 16);
                const float inv2 = inv * inv;
                q = fmaf(gj * g1j, inv2, q);
                invq[jr] = inv;
                if (pc == jr) myinv = inv;
                if (jr == 3 && gl == 0) *reinterpret_cast<float4*>(ib + 4 * ja) = make_float4(invq[0], invq[1], invq[2], invq[3]);
                __syncwarp();
                float nm[NB + 4], cv[NB];
                const float ninv2 = -inv2;
                static_for<t0, NB / 4 + 1>([&](auto tc) {
                    constexpr int t = decltype(tc)::value;
                    const float4 m4 = *reinterpret_cast<const float4*>(cbw + pr * RS + 4 * t);
                    fmul2_bcast(nm[4 * t], nm[4 * t + 1], m4.x, m4.y, ninv2);
                    fmul2_bcast(nm[4 * t + 2], nm[4 * t + 3], m4.z, m4.w, ninv2);
                    if constexpr (t < NB / 4) {
                        const float4 c4 = *reinterpret_cast<const float4*>(cbw + pc * RS + 4 * t);
                        cv[4 * t] = c4.x; cv[4 * t + 1] = c4.y; cv[4 * t + 2] = c4.z; cv[4 * t + 3] = c4.w;
                    }
                });
                static_for<ja, NB + 1>([&](auto ac) {
                    constexpr int a = decltype(ac)::value;
                    constexpr int bmax = a < NB ? a : NB - 1;
                    static_for<ja / 2, bmax / 2 + 1>([&](auto hc) {
                        constexpr int b0 = 2 * decltype(hc)::value;
                        // b0 may be ja - 1: the published column is zero for finished rows, so that half is a no-op
                        if constexpr (b0 + 1 <= bmax) ffma2_bcast(A[a][b0], A[a][b0 + 1], nm[a], cv[b0], cv[b0 + 1]);
                        else A[a][b0] = fmaf(nm[a], cv[b0], A[a][b0]);
                    });
                });
                if constexpr (jr == 3) {
#pragma unroll
                    for (int a = ja; a <= NB; ++a) A[a][ja] *= myinv;
                }
            });
            __syncwarp();
            const float hld = 0.5f * (float)VMP_LOG_2 * hl2;
            bad |= !(fabsf(hld) < CUDART_INF_F);
            const float score = scl[0] - 0.5f * q + 0.5f * scl[1] - hld;

            // ---------------- phase 3: samples, ELBO terms
            const uint64_t pair = (uint64_t)n * K + k;
            float snum = 0.f, sden = 0.f;
            for (int s = 0; s < S; ++s) {
                float e2;
                {
                    float4 e4;
                    if (p.noise != nullptr) {
                        e4.x = p.noise[(pair * D + 4 * gl + 0) * (uint64_t)S + s];
                        e4.y = p.noise[(pair * D + 4 * gl + 1) * (uint64_t)S + s];
                        e4.z = p.noise[(pair * D + 4 * gl + 2) * (uint64_t)S + s];
                        e4.w = p.noise[(pair * D + 4 * gl + 3) * (uint64_t)S + s];
                    } else {
                        e4 = philox_normal4(p.seed, pair, (uint32_t)s, (uint32_t)gl);
                    }
                    __syncwarp();
                    epsres[0 * RS + gl] = e4.x;
                    epsres[1 * RS + gl] = e4.y;
                    epsres[2 * RS + gl] = e4.z;
                    epsres[3 * RS + gl] = e4.w;
                    e2 = fmaf(e4.x, e4.x, fmaf(e4.y, e4.y, fmaf(e4.z, e4.z, e4.w * e4.w)));
                    __syncwarp();
                }
                float wpart[NB], xs[NB];
#pragma unroll
                for (int t = 0; t < NB / 4; ++t) {
                    const float4 v4 = *reinterpret_cast<const float4*>(epsres + pc * RS + 4 * t);
                    wpart[4 * t] = (pr == 0) ? v4.x - A[NB][4 * t] : 0.f;
                    wpart[4 * t + 1] = (pr == 0) ? v4.y - A[NB][4 * t + 1] : 0.f;
                    wpart[4 * t + 2] = (pr == 0) ? v4.z - A[NB][4 * t + 2] : 0.f;
                    wpart[4 * t + 3] = (pr == 0) ? v4.w - A[NB][4 * t + 3] : 0.f;
                }
                static_for_down<NB>([&](auto bc) {
                    constexpr int bb = decltype(bc)::value;
                    float t = wpart[bb];
                    t += __shfl_xor_sync(FULL, t, 4, 16);
                    t += __shfl_xor_sync(FULL, t, 8, 16);
                    const float4 ig4 = *reinterpret_cast<const float4*>(ib + 4 * bb);
                    const float ig[4] = {ig4.x, ig4.y, ig4.z, ig4.w};
                    float yb[4];
                    static_for_down<4>([&](auto ccc) {
                        constexpr int cc = decltype(ccc)::value;
                        yb[cc] = __shfl_sync(FULL, t, cc, 16) * ig[cc];
                        if constexpr (cc > 0) {
                            const float lv = __shfl_sync(FULL, A[bb][bb], cc * 4 + pc, 16);   // L[4bb+cc][4bb+pc]
                            if (pc < cc) t = fmaf(-lv, yb[cc], t);
                        }
                    });
                    const float ymine = (pr == 0) ? yb[0] : (pr == 1) ? yb[1] : (pr == 2) ? yb[2] : yb[3];
                    xs[bb] = (pc == 0) ? yb[0] : (pc == 1) ? yb[1] : (pc == 2) ? yb[2] : yb[3];
                    const float nym = -ymine;
                    static_for<0, (bb + 1) / 2>([&](auto hc) {
                        constexpr int b0 = 2 * decltype(hc)::value;
                        if constexpr (b0 + 1 < bb) ffma2_bcast(wpart[b0], wpart[b0 + 1], nym, A[bb][b0], A[bb][b0 + 1]);
                        else if constexpr (b0 < bb) wpart[b0] = fmaf(nym, A[bb][b0], wpart[b0]);
                    });
                });
                // x = mu1 + y (column class pc, replicated over pr)
#pragma unroll
                for (int t = 0; t < NB / 4; ++t) {
                    const float4 m4 = *reinterpret_cast<const float4*>(mures + pc * RS + 4 * t);
                    xs[4 * t] += m4.x; xs[4 * t + 1] += m4.y; xs[4 * t + 2] += m4.z; xs[4 * t + 3] += m4.w;
                }
                if (p.x_k_samples != nullptr && active && pr == 0) {
#pragma unroll
                    for (int b = 0; b < NB; ++b) p.x_k_samples[((pair * S) + s) * (uint64_t)D + 4 * b + pc] = xs[b];
                }
                if (s == 0) {
                    // online Gumbel-max draw of z_n (tf.multinomial GPU algorithm): keep the running arg-max and its sample
                    const float u = p.gum_u != nullptr ? p.gum_u[pair] : philox_uniform_pair(p.seed, pair);
                    const float cand = score + gumbel_from_uniform<float>(u);
                    if (cand > best) {
                        best = cand;
                        zbest = k;
                        if (pr == 0) {
#pragma unroll
                            for (int t = 0; t < NB / 4; ++t)
                                *reinterpret_cast<float4*>(xbres + pc * RS + 4 * t) =
                                    make_float4(xs[4 * t], xs[4 * t + 1], xs[4 * t + 2], xs[4 * t + 3]);
                        }
                    }
                }
                // |W (x - m)|^2
#pragma unroll
                for (int t = 0; t < NB / 4; ++t) {
                    const float4 m4 = *reinterpret_cast<const float4*>(mthres + pc * RS + 4 * t);
                    xs[4 * t] -= m4.x; xs[4 * t + 1] -= m4.y; xs[4 * t + 2] -= m4.z; xs[4 * t + 3] -= m4.w;
                }
                float q2 = 0.f;
                static_for<0, NB>([&](auto ac) {
                    constexpr int a = decltype(ac)::value;
                    float ta = 0.f;
                    static_for<0, a / 4 + 1>([&](auto tc) {
                        constexpr int t = decltype(tc)::value;
                        const float4 w4 = WL[(G::chbase(a) + t) * 16];
                        const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (4 * t + i <= a) ta = fmaf(wv[i], xs[4 * t + i], ta);
                    });
                    ta += __shfl_xor_sync(FULL, ta, 1, 16);
                    ta += __shfl_xor_sync(FULL, ta, 2, 16);
                    q2 = fmaf(ta, ta, q2);
                });
                q2 += __shfl_xor_sync(FULL, q2, 4, 16);
                q2 += __shfl_xor_sync(FULL, q2, 8, 16);
                e2 = group_sum<16>(e2);
                snum += -0.5f * e2;
                sden += (p.den_mode == VMP_DEN_GAUSS) ? (scl[2] - 0.5f * q2)
                                                      : (scl[2] - 0.5f * (scl[3] + (float)D) * log1pf(q2 / scl[3]));
            }
            if (gl == 0) {
                kst[3 * k + 0] = score;
                kst[3 * k + 1] = snum / (float)S + hld - 0.5f * (float)VMP_LOG_2PI * (float)D;
                kst[3 * k + 2] = sden / (float)S;
            }
        }

        // ---------------- per-point epilogue: log-sum-exp over K, log_r, ELBO partials, selected sample
        __syncwarp();
        float mx = -CUDART_INF_F;
        for (int k = gl; k < K; k += 16) mx = fmaxf(mx, kst[3 * k]);
        mx = group_max<16>(mx);
        double se = 0.0;
        for (int k = gl; k < K; k += 16) se += (double)expf(kst[3 * k] - mx);
        se = group_sum_d<16>(se);
        const float lse = mx + (float)log(se);
        double en = 0.0, ed = 0.0;
        if (active) {
            for (int k = gl; k < K; k += 16) {
                const float lr = kst[3 * k] - lse;
                p.log_r[n * K + k] = lr;
                const double r = (double)expf(lr);
                en += r * ((double)kst[3 * k + 1] + (double)lr);
                ed += r * (double)kst[3 * k + 2];
            }
            if (p.x_sample != nullptr) {
#pragma unroll
                for (int i = 0; i < D / 16; ++i) {
                    const int e = i * 16 + gl;
                    p.x_sample[n * D + e] = xbres[(e & 3) * RS + (e >> 2)];
                }
            }
            if (p.z != nullptr && gl == 0) p.z[n] = zbest;
        }
        en = group_sum_d<16>(en);
        ed = group_sum_d<16>(ed);
        if (gl == 0 && active) {
            atomicAdd(&cta_acc[0], en);
            atomicAdd(&cta_acc[1], ed);
            if (bad) atomicAdd(&cta_acc[3], 1.0);
        }
        __syncthreads();
    }
    __syncthreads();
    if (tid == 0) {
        atomicAdd(p.elbo_acc + 0, cta_acc[0]);
        atomicAdd(p.elbo_acc + 1, cta_acc[1]);
        atomicAdd(p.elbo_acc + 2, cta_acc[0] - cta_acc[1]);
        if (cta_acc[3] != 0.0) atomicAdd(p.elbo_acc + 3, cta_acc[3]);
    }
}

// (phi_rec, theta_rec) -> lane-major staged records of the 2-D engine
__global__ void pack_fast2d_records_kernel(int K, const float* __restrict__ phi_rec,
                                           const float* __restrict__ theta_rec, float* __restrict__ out) {
    using G = Fast2dGeom<64>;
    constexpr int D = 64, NB = G::NB, RS = G::RS, MAT = G::MAT, REC = G::REC;
    const int k = blockIdx.x;
    const float* prc = phi_rec + (size_t)k * phi_record_len(D);
    const float* trc = theta_rec + (size_t)k * theta_record_len(D);
    float* o = out + (size_t)k * REC;
    for (int e = threadIdx.x; e < MAT; e += blockDim.x) {
        const int i = e & 3, ln = (e >> 2) & 15, ch = e >> 6;
        int a = 0;
        while (a + 1 < NB && G::chbase(a + 1) <= ch) ++a;
        const int t = ch - G::chbase(a), b = 4 * t + i;
        const int pr = ln >> 2, pc = ln & 3, row = 4 * a + pr, col = 4 * b + pc;
        const bool ok = b <= a && row >= col;
        o[e] = ok ? prc[row * D + col] : 0.f;
        o[MAT + e] = ok ? trc[row * D + col] : 0.f;
    }
    for (int e = threadIdx.x; e < 4 * RS; e += blockDim.x) {
        const int r = e / RS, t = e - r * RS;
        o[2 * MAT + e] = t < NB ? prc[D * D + 4 * t + r] : 0.f;              // mu2
        o[2 * MAT + 4 * RS + e] = t < NB ? trc[D * D + 4 * t + r] : 0.f;     // m_theta
    }
    if (threadIdx.x < 8) {
        float v = 0.f;
        if (threadIdx.x == 0) v = prc[D * D + 2 * D];        // log pi
        if (threadIdx.x == 1) v = prc[D * D + 2 * D + 1];    // logdet P2
        if (threadIdx.x == 2) v = trc[D * D + D];            // cden
        if (threadIdx.x == 3) v = trc[D * D + D + 1];        // nu
        o[2 * MAT + 8 * RS + threadIdx.x] = v;
    }
}

void launch_pack_fast2d_records(int K, const float* phi_rec, const float* theta_rec, float* out, cudaStream_t st) {
    pack_fast2d_records_kernel<<<K, 256, 0, st>>>(K, phi_rec, theta_rec, out);
}

int launch_fast2d_64(const FastParams& p0, cudaStream_t st) {
    constexpr int WARPS = 8, PPC = WARPS * 2;
    FastParams p = p0;
    p.ntiles = (p.N + PPC - 1) / PPC;
    const size_t smem = fast2d_smem_bytes(p.K);
    auto kern = local_step_fast2d_kernel<64, WARPS>;
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
#endif  // VMP_FAST2D_IMPL

}  // namespace vmp
