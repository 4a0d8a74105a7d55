// fast_pack.cu — (phi_rec, theta_rec) of prepare.cu -> the padded, staged per-component records the group engine
// pulls into shared memory with one bulk copy per component: P2[D][D+4] | W[D][D+4] | mu2[D] | m[D] | 8 scalars.
#include "local_step_fast.cuh"

namespace vmp {

__global__ void pack_fast_records_kernel(int K, int Dr, int D, const float* __restrict__ phi_rec,
                                         const float* __restrict__ theta_rec, float* __restrict__ out) {
    // Dr: the caller's latent dimension (layout of phi_rec / theta_rec); D >= Dr: the engine dimension.  Rows / columns
    // Dr..D-1 extend P2 by an identity block (P~ = P2 + diag(p1) stays SPD, log-det and solves are unchanged) and W,
    // mu2, m by zeros.
    const int k = blockIdx.x;
    const int LD = D + 4, REC = fast_rec_len(D);
    const float* pr = phi_rec + (size_t)k * phi_record_len(Dr);
    const float* tr = theta_rec + (size_t)k * theta_record_len(Dr);
    float* o = out + (size_t)k * REC;
    for (int e = threadIdx.x; e < D * LD; e += blockDim.x) {
        const int i = e / LD, c = e - i * LD;
        const bool in = i < Dr && c < Dr;
        o[e] = in ? pr[i * Dr + c] : ((i == c && c < D) ? 1.f : 0.f);
        o[D * LD + e] = in ? tr[i * Dr + c] : 0.f;
    }
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        o[2 * D * LD + i] = i < Dr ? pr[Dr * Dr + i] : 0.f;           // mu2
        o[2 * D * LD + D + i] = i < Dr ? tr[Dr * Dr + i] : 0.f;       // m_theta
    }
    if (threadIdx.x < 8) {
        float v = 0.f;
        if (threadIdx.x == 0) v = pr[Dr * Dr + 2 * Dr];        // log pi
        if (threadIdx.x == 1) v = pr[Dr * Dr + 2 * Dr + 1];    // logdet P2
        if (threadIdx.x == 2) v = tr[Dr * Dr + Dr];            // cden
        if (threadIdx.x == 3) v = tr[Dr * Dr + Dr + 1];        // nu
        o[2 * D * LD + 2 * D + threadIdx.x] = v;
    }
}

void launch_pack_fast_records(int K, int Dr, int De, const float* phi_rec, const float* theta_rec, float* out,
                              cudaStream_t st) {
    pack_fast_records_kernel<<<K, 256, 0, st>>>(K, Dr, De, phi_rec, theta_rec, out);
}

}  // namespace vmp
