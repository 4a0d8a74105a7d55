/* vmp_svae.h — C-ABI of libvmp_svae.so: the B200 (sm_100a) implementation of the local VMP step +
 * natural-gradient global update of emtiyaz/vmp-for-svae.
 *
 * The reference has no FFI of its own: its hot path is the Python function surface of
 * models/svae.py, models/gmm.py, models/smm.py and distributions/{gaussian,niw,dirichlet,student_t}.py,
 * executed by TensorFlow 1.3 library kernels.  Each entry point below names the reference
 * function(s) (file:line) whose arithmetic it replaces.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer to caller-owned, contiguous, row-major memory; the library
 *    never allocates device memory and keeps no global state;
 *  - `stream` is a cudaStream_t passed as void*; all calls are asynchronous on it, no hidden syncs;
 *  - `_f32` entry points take float buffers, `_f64` entry points double buffers (the fp64 build of the
 *    north-star tolerance clause); integer outputs are int32; ELBO / sufficient-statistic accumulators
 *    are always double;
 *  - return value: 0 ok; <0 invalid argument (VMP_E_*); >0 a cudaError_t from the launch;
 *  - N points, K mixture components, D latent dimension (1..64), S samples per (point, component).
 */
#ifndef VMP_SVAE_H
#define VMP_SVAE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VMP_OK            0
#define VMP_E_BADARG     -1   /* null pointer / non-positive size */
#define VMP_E_BADDIM     -2   /* D outside 1..VMP_MAX_D */
#define VMP_E_BADMODE    -3

#define VMP_MAX_D        64

/* denominators of the ELBO regulariser */
#define VMP_DEN_GAUSS     0   /* log N(x | E_theta)   : svae.compute_elbo      (svae.py:199-262) */
#define VMP_DEN_STUDENT   1   /* log St(x | mu,Sigma,nu): svae.compute_elbo_smm (svae.py:265-322) */

int vmp_version(void);

/* Number of scalars in one per-component record (see DESIGN.md "data layout"). */
int vmp_phi_record_len(int D);      /* P2[D*D] | mu2[D] | h2[D] | log_pi, logdetP2, 0, 0            */
int vmp_theta_record_len(int D);    /* W[D*D]  | m[D]   | cden, nu, E log pi, logdet P               */
int vmp_stats_len(int D);           /* per component: N_k, W_k, sum w x [D], sum w x x^T [D*D]       */

/* ---- per-step prologues (K-sized) -------------------------------------------------------------------
 * phi_gmm -> per-component records.  Replaces svae.unpack_recognition_gmm (svae.py:342-358):
 * L = tril(L_raw) with softplus diagonal, P2 = L L^T, pi = softmax(pi_raw); additionally mu2 = P2^-1 eta1,
 * logdet P2.  eta1_phi2[K,D], L_raw[K,D,D], pi_raw[K] -> rec[K, vmp_phi_record_len(D)].               */
int vmp_phi_prepare_f32(int K, int D, const float* eta1_phi2, const float* L_raw, const float* pi_raw,
                        float* rec, void* stream);
int vmp_phi_prepare_f64(int K, int D, const double* eta1_phi2, const double* L_raw, const double* pi_raw,
                        double* rec, void* stream);

/* theta (Dirichlet+NIW natural parameters) -> records for the ELBO denominator.  Replaces
 * niw.natural_to_standard + niw.expected_values + gaussian.standard_to_natural + dirichlet.expected_log_pi
 * as called from svae.compute_elbo (svae.py:204-208; niw.py:8-17,33-43; gaussian.py:11-19; dirichlet.py:8-12).
 * alpha[K], A[K,D,D], b[K,D], beta[K], v_hat[K] -> rec[K, vmp_theta_record_len(D)].                   */
int vmp_theta_prepare_gauss_f32(int K, int D, const float* alpha, const float* A, const float* b,
                                const float* beta, const float* v_hat, float* rec, void* stream);
int vmp_theta_prepare_gauss_f64(int K, int D, const double* alpha, const double* A, const double* b,
                                const double* beta, const double* v_hat, double* rec, void* stream);
/* SMM variant: theta = (alpha_nat[K], mu[K,D], L_raw[K,D,D], dof[K]).  Replaces svae.unpack_smm
 * (svae.py:361-373) + the per-component constants of student_t._logprob_full_scale (student_t.py:7-39). */
int vmp_theta_prepare_student_f32(int K, int D, const float* alpha, const float* mu, const float* L_raw,
                                  const float* dof, float* rec, void* stream);
int vmp_theta_prepare_student_f64(int K, int D, const double* alpha, const double* mu, const double* L_raw,
                                  const double* dof, double* rec, void* stream);

/* ---- the fused local step ---------------------------------------------------------------------------
 * Replaces svae.e_step (svae.py:14-47: compute_log_z_given_y 50-92, sample_x_per_comp 95-119),
 * svae.subsample_x(...)[:,0,:] (122-151, 514) and the regulariser of svae.compute_elbo /
 * compute_elbo_smm (234-260 / 294-320) incl. gaussian.log_probability_nat (gaussian.py:30-71),
 * log_probability_nat_per_samp (74-105), student_t.log_probability_per_samp (student_t.py:59-61).
 *
 * in : eta1[N,D], eta2_diag[N,D] (<0)  encoder natural parameters
 *      phi_rec, theta_rec               from the prologues; den_mode VMP_DEN_*
 *      noise[N,K,D,S] or NULL           injected raw noise (svae.py:113-114 layout); NULL -> Philox(seed)
 *      gumbel_u[N,K] or NULL            injected uniforms of the categorical draw z_n = argmax_k(log r_nk -
 *                                       log(-log u_nk)) (tf.multinomial's GPU algorithm); NULL -> Philox(seed)
 *      point_offset                     global index of point 0 of this call: the in-kernel Philox streams are keyed by
 *                                       the GLOBAL pair index (point_offset + n) * K + k, so a batch sharded across
 *                                       ranks (data.py:174-175) or processed in host chunks draws exactly what one call
 *                                       over the whole batch draws.  Ignored for injected noise / gumbel_u.
 *      x_in[N,K,S,D] or NULL            evaluate these samples instead of drawing (svae.compute_elbo called with
 *                                       caller-supplied x_k_samps); eps is recovered as a + L^T (x - mu1)
 * out: log_r[N,K]                       normalised log q(z|y)         (may not be NULL)
 *      x_sample[N,D], z[N]              the selected sample x[n, z_n, 0] and z_n (either may be NULL)
 *      x_k_samples[N,K,S,D] or NULL     all samples (only materialise at small sizes)
 *      elbo_acc[4] (double, ACCUMULATED into; caller zeroes): sum r*mean_s(log num), sum r*mean_s(log den),
 *                                       regulariser = mean_s sum r (num - den), number of non-PD pivots   */
int vmp_svae_local_step_f32(int64_t N, int K, int D, int S, const float* eta1, const float* eta2_diag,
                            const float* phi_rec, const float* theta_rec, int den_mode,
                            const float* noise, const float* gumbel_u, uint64_t seed, int64_t point_offset,
                            const float* x_in, float* log_r, float* x_sample, int32_t* z, float* x_k_samples,
                            double* elbo_acc, void* workspace, size_t workspace_bytes, void* stream);
int vmp_svae_local_step_f64(int64_t N, int K, int D, int S, const double* eta1, const double* eta2_diag,
                            const double* phi_rec, const double* theta_rec, int den_mode,
                            const double* noise, const double* gumbel_u, uint64_t seed, int64_t point_offset,
                            const double* x_in, double* log_r, double* x_sample, int32_t* z, double* x_k_samples,
                            double* elbo_acc, void* workspace, size_t workspace_bytes, void* stream);
/* Caller-provided scratch for the step: the staged per-component records of the fp32 group engine (9 <= D <= 64 runs
 * in an engine of dimension 16 / 32 / 64).  Without it (NULL / too small), for D <= 8, for fp64 and for caller-supplied
 * samples (x_in) the thread-per-pair kernels run instead.  No environment variable selects anything.     */
size_t vmp_svae_local_step_workspace_bytes(int K, int D);

/* ---- the whole step in one launch (launch-bound configurations: BASELINE C1 / C2) --------------------------------------
 * phi / theta prologues + local step + selection + statistics + natural-gradient update as ONE kernel on one thread-block
 * cluster (csrc/small_step.cu); replaces the same reference lines as vmp_phi_prepare + vmp_theta_prepare_* +
 * vmp_svae_local_step + vmp_suffstats + vmp_ng_update (svae.py:14-176, 199-322, 342-403).  Single GPU; D <= 8, K <= 32,
 * N*K <= 8192 (vmp_svae_small_step_supported).  theta / prior / theta_out are HOST arrays of device pointers:
 *   den_mode VMP_DEN_GAUSS   : theta = {alpha, A, b, beta, v_hat}; VMP_DEN_STUDENT: theta = {alpha, mu, L_raw, dof}
 *   prior = {alpha, A, b, beta, v_hat} (only prior[0] when only_alpha); theta_out = the tensors updated in place
 *   (theta itself for the GMM; {alpha} with only_alpha != 0 for the SMM, svae.m_step_smm 179-196).
 * stats[K, vmp_stats_len(D)] and elbo_acc[4] are OVERWRITTEN (not accumulated).  Other arguments as vmp_svae_local_step_*.  */
int vmp_svae_small_step_supported(int64_t N, int K, int D);
int vmp_svae_small_step_f32(int64_t N, int K, int D, int S, int den_mode, int only_alpha, const float* eta1,
                            const float* eta2_diag, const float* eta1_phi2, const float* L_raw, const float* pi_raw,
                            const float* const* theta, const float* const* prior, float* const* theta_out, double rho,
                            const double* rho_dev, const float* noise, const float* gumbel_u, uint64_t seed,
                            int64_t point_offset, float* log_r, float* x_sample, int32_t* z, float* x_k_samples, double* stats,
                            double* elbo_acc, void* stream);
int vmp_svae_small_step_f64(int64_t N, int K, int D, int S, int den_mode, int only_alpha, const double* eta1,
                            const double* eta2_diag, const double* eta1_phi2, const double* L_raw, const double* pi_raw,
                            const double* const* theta, const double* const* prior, double* const* theta_out, double rho,
                            const double* rho_dev, const double* noise, const double* gumbel_u, uint64_t seed,
                            int64_t point_offset, double* log_r, double* x_sample, int32_t* z, double* x_k_samples,
                            double* stats, double* elbo_acc, void* stream);

/* The same call with its arguments in one struct (dtype: 0 = f32, 1 = f64; pointers as in vmp_svae_small_step_*): a host
 * language binding fills the struct ONCE per set of buffers and only rewrites rho / seed per step — at the launch-bound shapes
 * the per-argument marshalling of a 32-argument FFI call costs as much as the kernel.                                     */
typedef struct VmpSmallStepArgs {
    int64_t N;
    int32_t K, D, S, den_mode, only_alpha, dtype;
    const void *eta1, *eta2_diag, *eta1_phi2, *L_raw, *pi_raw;
    const void* theta[5];
    const void* prior[5];
    void* theta_out[5];
    double rho;
    const double* rho_dev;
    const void *noise, *gumbel_u;
    uint64_t seed;
    int64_t point_offset;
    void *log_r, *x_sample;
    int32_t* z;
    void* x_k_samples;
    double *stats, *elbo_acc;
    void* stream;
} VmpSmallStepArgs;
int vmp_svae_small_step_packed(const VmpSmallStepArgs* args);

/* ---- reverse pass of the fused local step ---------------------------------------------------------------
 * Replaces what TF's autodiff builds for opt.compute_gradients(-elbo) (experiments.py:232) through svae.e_step
 * (svae.py:39-100), the per-component sampling (svae.py:103-123) and the regulariser of compute_elbo / compute_elbo_smm
 * (svae.py:199-322); theta is a constant there (tf.stop_gradient, svae.py:211-214).  Computes the gradients of
 *      sum(gx * x_k_samples) + sum(glr * log_r) + greg * regulariser          (regulariser == elbo_acc[2] of the step)
 * w.r.t. eta1[N,D], eta2_diag[N,D] and the raw phi_gmm (eta1_phi2[K,D], L_raw[K,D,D], pi_raw[K]).  The caller passes
 * the same noise/seed and the log_r of the forward call; gx[N,K,S,D], glr[N,K] are the upstream gradients; greg_dev
 * (device scalar, may be NULL) overrides greg so that no host synchronisation is needed (CUDA-graph capture).
 * theta_rec_bar[K, vmp_theta_record_len(D)] (may be NULL) receives the gradient w.r.t. the theta record (W lower | m |
 * cden, zeros): compute_elbo_smm trains mu_k, L_k of the Student-t components by gradient (svae.py:265-322 has no
 * stop_gradient on them; experiments.py:154-174).  The in-kernel noise is keyed with point_offset 0.  D <= 16 and K <= 256: thread-per-pair kernel; larger D (<= 64) or K: one CTA
 * per point with the pair's matrices in shared memory (block-cooperative Cholesky / triangular inverse / Murray reverse).  workspace: vmp_svae_local_step_bwd_workspace_bytes(K, D) bytes.            */
#define VMP_BWD_MAX_D 64
size_t vmp_svae_local_step_bwd_workspace_bytes(int K, int D);
int vmp_svae_local_step_bwd_f32(int64_t N, int K, int D, int S, const float* eta1, const float* eta2_diag,
                                const float* eta1_phi2, const float* L_raw, const float* pi_raw, const float* phi_rec,
                                const float* theta_rec, int den_mode, const float* noise, uint64_t seed,
                                const float* log_r, const float* gx, const float* glr, double greg, const float* greg_dev,
                                float* eta1_bar,
                                float* eta2_diag_bar, float* eta1_phi2_bar, float* L_raw_bar, float* pi_raw_bar,
                                float* theta_rec_bar, void* workspace, size_t workspace_bytes, void* stream);
int vmp_svae_local_step_bwd_f64(int64_t N, int K, int D, int S, const double* eta1, const double* eta2_diag,
                                const double* eta1_phi2, const double* L_raw, const double* pi_raw,
                                const double* phi_rec, const double* theta_rec, int den_mode, const double* noise,
                                uint64_t seed, const double* log_r, const double* gx, const double* glr, double greg,
                                const double* greg_dev, double* eta1_bar, double* eta2_diag_bar, double* eta1_phi2_bar, double* L_raw_bar,
                                double* pi_raw_bar, double* theta_rec_bar, void* workspace, size_t workspace_bytes,
                                void* stream);

/* The noise the in-kernel generator uses for a given seed, written in the reference layout
 * (tests: injected-noise path == in-kernel path).  noise[N,K,D,S], u[N,K] (either may be NULL).
 * Definition of the stream: block(pair, s, q) = Philox4x32-10(counter = (pair_lo, pair_hi, s, q), key = seed) with
 * pair = (point_offset + n) * K + k; the four normals of dims 4q..4q+3 of sample s are two Box-Muller pairs of the
 * words' top 23 bits (bin centres (j + 1/2) / 2^23; radius and direction evaluated with lg2 / sqrt / sin / cos .approx);
 * u[n,k] is built from the 9 unused low bits of three words of block(pair, 0, 0).                                 */
int vmp_fill_noise_f32(int64_t N, int K, int D, int S, uint64_t seed, int64_t point_offset, float* noise, float* u,
                       void* stream);
int vmp_fill_noise_f64(int64_t N, int K, int D, int S, uint64_t seed, int64_t point_offset, double* noise, double* u,
                       void* stream);

/* ---- responsibility-weighted sufficient statistics --------------------------------------------------
 * Replaces the reductions of gmm.m_step (gmm.py:25-46,201-227: update_Nk/xk/Sk) and smm.m_step
 * (smm.py:25-50,167-196) in additive form: stats[k] += [sum r, sum w, sum w x, sum w x x^T], w = r (GMM) or
 * r*u (SMM, u_nk != NULL).  x[N,D]; r[N,K] = responsibilities, or their logs when r_is_log != 0
 * (the SVAE path feeds log_r straight in: experiments.py:258-259 does tf.exp).  stats[K, vmp_stats_len(D)]
 * is double and ACCUMULATED into (zero it first; all-reduce it across ranks before the update).
 * Which kernel runs is a function of the arguments alone.  fp32: D = 64 with even K and GMM weights -> tcgen05 contraction
 * (suffstats_tc.cu); D in {16, 32} with K % 4 == 0 -> warp-level mma.sync (suffstats_mma.cu); D <= 8, K <= 32, plain r ->
 * the sweep statistics of mixture_sweep.cu (mma.sync at D = 8, K % 4 == 0); everything else and fp64 -> FP32 / FP64 FMA
 * kernels (suffstats.cu).  The tensor-core paths feed split-tf32 operands (hi*hi + lo*hi + hi*lo) and agree with an fp64
 * contraction to <= 3e-6 of each block's magnitude; the second-moment block is exactly symmetric on every path.        */
int vmp_suffstats_f32(int64_t N, int K, int D, const float* x, const float* r, int r_is_log,
                      const float* u_nk, double* stats, void* stream);
int vmp_suffstats_f64(int64_t N, int K, int D, const double* x, const double* r, int r_is_log,
                      const double* u_nk, double* stats, void* stream);

/* Statistics + natural-gradient update in ONE launch (single-GPU path; north_star: "the NIW/Dirichlet natural-gradient
 * step is fused into the tail of that reduction"): as vmp_suffstats_*, and the CTA that finishes last applies
 * theta <- (1-rho) theta + rho (prior + stats terms) exactly as vmp_ng_update_* (svae.py:154-176, 376-403).
 * counter: one device uint32, zero before the first call (the kernel resets it); stats is accumulated into as usual and
 * holds the statistics afterwards.  With several ranks use vmp_suffstats_* + all-reduce + vmp_ng_update_*.            */
int vmp_suffstats_update_f32(int64_t N, int K, int D, const float* x, const float* r, int r_is_log, const float* u_nk,
                             double* stats, unsigned int* counter, double rho, const double* rho_dev, int only_alpha,
                             const float* p_alpha, const float* p_A, const float* p_b, const float* p_beta,
                             const float* p_vhat, float* alpha, float* A, float* b, float* beta, float* v_hat, void* stream);
int vmp_suffstats_update_f64(int64_t N, int K, int D, const double* x, const double* r, int r_is_log, const double* u_nk,
                             double* stats, unsigned int* counter, double rho, const double* rho_dev, int only_alpha,
                             const double* p_alpha, const double* p_A, const double* p_b, const double* p_beta,
                             const double* p_vhat, double* alpha, double* A, double* b, double* beta, double* v_hat,
                             void* stream);

/* ---- natural-gradient (CVI) update ------------------------------------------------------------------
 * Replaces svae.m_step (svae.py:154-176) + svae.update_gmm_params (376-403):
 * theta* = prior + [N_k, sum r x x^T, sum r x, N_k, N_k + 1]; theta <- (1-rho) theta + rho theta*, in place.
 * theta_star_* (optional, may be NULL) receive theta*.  With only_alpha != 0 only alpha is touched
 * (svae.m_step_smm, svae.py:179-196).  rho_dev (device pointer, may be NULL) overrides rho: a device-resident
 * step size lets a captured CUDA graph follow the decaying CVI schedule (experiments.py:143-147).          */
int vmp_ng_update_f32(int K, int D, const double* stats, double rho, const double* rho_dev, int only_alpha,
                      const float* p_alpha, const float* p_A, const float* p_b, const float* p_beta, const float* p_vhat,
                      float* alpha, float* A, float* b, float* beta, float* v_hat,
                      float* s_alpha, float* s_A, float* s_b, float* s_beta, float* s_vhat, void* stream);
int vmp_ng_update_f64(int K, int D, const double* stats, double rho, const double* rho_dev, int only_alpha,
                      const double* p_alpha, const double* p_A, const double* p_b, const double* p_beta, const double* p_vhat,
                      double* alpha, double* A, double* b, double* beta, double* v_hat,
                      double* s_alpha, double* s_A, double* s_b, double* s_beta, double* s_vhat, void* stream);

/* ---- standalone mixture VB-EM -----------------------------------------------------------------------
 * M-step in standard parameters from accumulated statistics.  Replaces gmm.update_* (gmm.py:25-81) when
 * is_smm == 0 (NaN guards of 36,46; v_k = v_0 + N_k + 1) and smm.update_* (smm.py:25-85) when is_smm != 0
 * (1e-20 eps; v_k = v_0 + N_k).  Outputs alpha_k[K], beta_k[K], m_k[K,D], C_k[K,D,D], v_k[K], x_k[K,D], S_k[K,D,D]. */
int vmp_mixture_mstep_f32(int K, int D, int is_smm, const double* stats,
                          const float* alpha_0, const float* beta_0, const float* m_0, const float* C_0, const float* v_0,
                          float* alpha_k, float* beta_k, float* m_k, float* C_k, float* v_k, float* x_k, float* S_k,
                          void* stream);
int vmp_mixture_mstep_f64(int K, int D, int is_smm, const double* stats,
                          const double* alpha_0, const double* beta_0, const double* m_0, const double* C_0, const double* v_0,
                          double* alpha_k, double* beta_k, double* m_k, double* C_k, double* v_k, double* x_k, double* S_k,
                          void* stream);

/* E-steps.  gmm.e_step / e_step_missing_data (gmm.py:154-198 with 84-151) when kappa_k == NULL;
 * smm.e_step (smm.py:140-164 with 88-137) otherwise.  x[N,D], alpha_k[K], beta_k[K], m_k[K,D], P_k[K,D,D],
 * v_k[K], missing_mask[N,D] (uint8, may be NULL; GMM only) -> r[N,K], u_out[N,K] (SMM only), pi[K] = exp(E log pi).
 * work[K]: caller-provided scratch for the per-component constants.                                    */
int vmp_mixture_estep_f32(int64_t N, int K, int D, const float* x, const float* alpha_k, const float* beta_k,
                          const float* m_k, const float* P_k, const float* v_k, const float* kappa_k,
                          const uint8_t* missing_mask, float* r, float* u_out, float* pi, float* work, void* stream);
int vmp_mixture_estep_f64(int64_t N, int K, int D, const double* x, const double* alpha_k, const double* beta_k,
                          const double* m_k, const double* P_k, const double* v_k, const double* kappa_k,
                          const uint8_t* missing_mask, double* r, double* u_out, double* pi, double* work,
                          void* stream);

/* Whole VB-EM sweeps: n_sweeps x [ m_step(r[,u]) -> P = inv(C) -> e_step -> (r[,u]) ], i.e. gmm.inference (gmm.py:230-269) /
 * smm.inference (smm.py:199-245) applied n_sweeps times to the state (the reference's driver loops gmm.py:377-379 and
 * smm.py __main__).  In: x[N,D]; the prior in standard parameters alpha_0[K], beta_0[K], m_0[K,D], C_0[K,D,D], v_0[K]
 * (niw.natural_to_standard / dirichlet.natural_to_standard of init_mm_params, gmm.py:246-252); kappa_k[K] (SMM, is_smm != 0);
 * state r[N,K] (and u[N,K] for the SMM), read as the input of the first M-step and overwritten with the state after the last
 * sweep.  Out (of the LAST sweep's M-step): alpha_k, beta_k, m_k, C_k, v_k, x_k, S_k, pi = exp(E log pi).
 * fp32 with D <= 8 and K <= 32 runs the fused kernels of mixture_sweep.cu (inside a multi-sweep call r and u stay on chip:
 * the e-step accumulates the next M-step's statistics; only the last sweep writes the state); other shapes run the general
 * kernels sweep by sweep.  workspace: vmp_mixture_fit_workspace_bytes(K, D) bytes of device scratch.                   */
size_t vmp_mixture_fit_workspace_bytes(int K, int D);
/* The phases of one sweep as separate calls (a multi-rank sweep all-reduces the statistics between them; vmp_mixture_fit is
 * their single-process composition):
 *   vmp_suffstats_*            statistics of the state (r[,u])
 *   vmp_mixture_prepare_*      K-sized: M-step in standard parameters from the (all-reduced) statistics + P_k = C_k^-1 + the
 *                              e-step constants; optional outputs P_k[K,D,D], cst[K] (inputs of vmp_mixture_estep_*) and
 *                              rec[K, vmp_mixture_record_len(D)] (input of vmp_mixture_estep_fused_f32; D <= 8);
 *                              stats_next (may be NULL) is zeroed for the next accumulation
 *   vmp_mixture_estep_fused_f32   fp32, D <= 8, K <= 32: e-step from the packed records; writes r[,u] when write_state != 0
 *                              and / or accumulates the statistics of the NEW state into stats_next (non-NULL)            */
int vmp_mixture_record_len(int D);
int vmp_mixture_prepare_f32(int K, int D, int is_smm, const double* stats, double* stats_next, const float* alpha_0,
                            const float* beta_0, const float* m_0, const float* C_0, const float* v_0, const float* kappa_k,
                            float* alpha_k, float* beta_k, float* m_k, float* C_k, float* v_k, float* x_k, float* S_k, float* pi,
                            float* P_k, float* cst, float* rec, void* stream);
int vmp_mixture_prepare_f64(int K, int D, int is_smm, const double* stats, double* stats_next, const double* alpha_0,
                            const double* beta_0, const double* m_0, const double* C_0, const double* v_0, const double* kappa_k,
                            double* alpha_k, double* beta_k, double* m_k, double* C_k, double* v_k, double* x_k, double* S_k,
                            double* pi, double* P_k, double* cst, float* rec, void* stream);
int vmp_mixture_estep_fused_f32(int64_t N, int K, int D, int is_smm, const float* x, const float* rec, float* r, float* u,
                                double* stats_next, int write_state, void* stream);
int vmp_mixture_fit_f32(int64_t N, int K, int D, int is_smm, int n_sweeps, const float* x, const float* alpha_0,
                        const float* beta_0, const float* m_0, const float* C_0, const float* v_0, const float* kappa_k, float* r,
                        float* u, float* alpha_k, float* beta_k, float* m_k, float* C_k, float* v_k, float* x_k, float* S_k,
                        float* pi, void* workspace, size_t workspace_bytes, void* stream);
int vmp_mixture_fit_f64(int64_t N, int K, int D, int is_smm, int n_sweeps, const double* x, const double* alpha_0,
                        const double* beta_0, const double* m_0, const double* C_0, const double* v_0, const double* kappa_k,
                        double* r, double* u, double* alpha_k, double* beta_k, double* m_k, double* C_k, double* v_k, double* x_k,
                        double* S_k, double* pi, void* workspace, size_t workspace_bytes, void* stream);

/* Batched SPD inverse + log-determinant of K matrices (tf.matrix_inverse at gmm.py:260 / smm.py:234,
 * helpers/tf_utils.logdet 25-49).  in[K,D,D] -> inv[K,D,D] (may be NULL), logdet[K] (may be NULL).      */
int vmp_spd_inverse_f32(int K, int D, const float* in, float* inv, float* logdet, void* stream);
int vmp_spd_inverse_f64(int K, int D, const double* in, double* inv, double* logdet, void* stream);

/* ---- ELBO terms outside the fused step --------------------------------------------------------------
 * Decoder-side weighted reductions.  mode 0: vae.expected_diagonal_gaussian_loglike, weighted branch
 * (vae.py:226-248): acc += sum_{n,k,s,d} w_nk [ (y-mean)^2/var + log(var+1e-8) ]; mode 1:
 * vae.expected_bernoulli_loglike (vae.py:175-198): acc += sum_{n,k,s,d} w_nk [ -log(1+exp(-logit*y)) ].
 * y[N,Dobs], means/out2[N,K,S,Dobs] (means may be NULL in mode 1), w[N,K]; acc is one double, accumulated. */
int vmp_decoder_loglike_f32(int64_t N, int K, int S, int Dobs, int mode, const float* y, const float* means,
                            const float* out2, const float* w, double* acc, void* stream);
int vmp_decoder_loglike_f64(int64_t N, int K, int S, int Dobs, int mode, const double* y, const double* means,
                            const double* out2, const double* w, double* acc, void* stream);

/* Reverse of the reduction above (what TF's autodiff builds for the neg_rec term of the ELBO, svae.py:220-223):
 * g_means = scale * d acc/d means, g_out2 = scale * d acc/d out2 [N,K,S,Dobs]; g_w = scale * d acc/d w [N,K] (may be
 * NULL).  mode 1 ignores means / g_means.                                                                */
int vmp_decoder_loglike_bwd_f32(int64_t N, int K, int S, int Dobs, int mode, const float* y, const float* means,
                                const float* out2, const float* w, double scale, float* g_means, float* g_out2,
                                float* g_w, void* stream);
int vmp_decoder_loglike_bwd_f64(int64_t N, int K, int S, int Dobs, int mode, const double* y, const double* means,
                                const double* out2, const double* w, double scale, double* g_means, double* g_out2,
                                double* g_w, void* stream);

/* Test-time metrics over the decoder outputs (losses.py): one pass yields, per (n,k),
 *   sq[n,k]  = 1/S sum_s sum_d m_d (target_nd - means_nksd)^2   -> losses.weighted_mse (9-40), imputation_mse (147-170)
 *   lse[n,k] = log sum_s exp(log_w_nks + sum_d m_d log p(y_nd | means_nksd, out2_nksd))
 *              -> losses.diagonal_gaussian_logprob (83-144, mode 0: out2 = variances), bernoulli_logprob (43-80, mode 1:
 *                 out2 = logits).  mask[N,Dobs] uint8 (1 = counted; NULL = all), target NULL = y, log_w_nks[N,K,S] NULL = 0;
 * sq / lse may be NULL.  The K-sized tail (weights, log-sum-exp over k, mean over n) is left to the caller.       */
int vmp_decoder_metrics_f32(int64_t N, int K, int S, int Dobs, int mode, const float* y, const float* target,
                            const float* means, const float* out2, const uint8_t* mask, const float* log_w_nks,
                            float* sq, float* lse, void* stream);
int vmp_decoder_metrics_f64(int64_t N, int K, int S, int Dobs, int mode, const double* y, const double* target,
                            const double* means, const double* out2, const uint8_t* mask, const double* log_w_nks,
                            double* sq, double* lse, void* stream);

/* General dense-natural-parameter Gaussian log-density (API surface of distributions/gaussian.py).
 * S == 0: gaussian.log_probability_nat (gaussian.py:30-71): x[N,D], eta1[N,K,D], eta2[N,K,D,D], log_w[K] or NULL
 *         -> out[N,K] normalised over K.   S >= 1: gaussian.log_probability_nat_per_samp (74-105):
 *         x[N,K,S,D] -> out[N,K,S].                                                                      */
int vmp_gaussian_logprob_nat_f32(int64_t N, int K, int S, int D, const float* x, const float* eta1,
                                 const float* eta2, const float* log_w, float* out, void* stream);
int vmp_gaussian_logprob_nat_f64(int64_t N, int K, int S, int D, const double* x, const double* eta1,
                                 const double* eta2, const double* log_w, double* out, void* stream);

/* Dense-natural-parameter sampling (API surface of svae.sample_x_per_comp, svae.py:95-119; the fused step never builds
 * these tensors): for B = N*K systems, P = -2 eta2[b] = L L^T (lower Cholesky, svae.py:111),
 * x[b,s,:] = P^-1 eta1[b] + L^-T noise[b,:,s]   (svae.py:115-118).  eta1[B,D], eta2[B,D,D], noise[B,D,S] -> x[B,S,D];
 * non_pd (device int, may be NULL) counts systems with a non-positive pivot.  Arithmetic in double.            */
int vmp_gaussian_sample_nat_f32(int64_t B, int D, int S, const float* eta1, const float* eta2, const float* noise,
                                float* x, int* non_pd, void* stream);
int vmp_gaussian_sample_nat_f64(int64_t B, int D, int S, const double* eta1, const double* eta2, const double* noise,
                                double* x, int* non_pd, void* stream);

/* FP32 pipe probe (measurement aid, not part of the reference surface): grid x 256 threads, each iters*128 FMAs as
 * scalar FFMA (packed == 0) or packed FFMA2 / fma.rn.f32x2 (packed != 0); out[grid*256] floats.  Timed by the caller. */
int vmp_fma_probe(int packed, int grid, int iters, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VMP_SVAE_H */
