// fast_pack.cu — (phi_rec, theta_rec) of prepare.cu -> the padded, staged per-component records the group engine
// pulls into shared memory with one bulk copy per component: P2[D][D+4] | W[D][D+4] | mu2[D] | m[D] | 8 scalars.
#include "local_step_fast.cuh"

namespace vmp {

__global__ void pack_fast_records_kernel(int K, int D, const float* __restrict__ phi_rec,
                                         const float* __restrict__ theta_rec, float* __restrict__ out) {
    const int k = blockIdx.x;
    const int LD = D + 4, REC = fast_rec_len(D);
    const float* pr = phi_rec + (size_t)k * phi_record_len(D);
    const float* tr = theta_rec + (size_t)k * theta_record_len(D);
    float* o = out + (size_t)k * REC;
    for (int e = threadIdx.x; e < D * LD; e += blockDim.x) {
        const int i = e / LD, c = e - i * LD;
        o[e] = c < D ? pr[i * D + c] : 0.f;
        o[D * LD + e] = c < D ? tr[i * D + c] : 0.f;
    }
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        o[2 * D * LD + i] = pr[D * D + i];           // mu2
        o[2 * D * LD + D + i] = tr[D * D + i];       // m_theta
    }
    if (threadIdx.x < 8) {
        float v = 0.f;
        if (threadIdx.x == 0) v = pr[D * D + 2 * D];        // log pi
        if (threadIdx.x == 1) v = pr[D * D + 2 * D + 1];    // logdet P2
        if (threadIdx.x == 2) v = tr[D * D + D];            // cden
        if (threadIdx.x == 3) v = tr[D * D + D + 1];        // nu
        o[2 * D * LD + 2 * D + threadIdx.x] = v;
    }
}

void launch_pack_fast_records(int K, int D, const float* phi_rec, const float* theta_rec, float* out, cudaStream_t st) {
    pack_fast_records_kernel<<<K, 256, 0, st>>>(K, D, phi_rec, theta_rec, out);
}

}  // namespace vmp
