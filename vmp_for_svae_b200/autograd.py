"""Differentiable fused local step: the piece of the reference's training graph between the encoder outputs and the
decoder input / ELBO regulariser, with the hand-written reverse pass (csrc/local_step_bwd.cu) behind
`torch.autograd.Function`.

Reference: experiments.py:208-232 builds  elbo = E log p(y|x) - regulariser  from svae.inference (svae.py:325-339) and
calls opt.compute_gradients(-elbo) on phi_enc / phi_gmm / decoder weights; TF differentiates through e_step
(svae.py:39-100), the per-component samples (svae.py:103-123) and compute_elbo (svae.py:199-262) with theta behind
tf.stop_gradient (svae.py:211-214).  Here those three are ONE forward kernel and ONE backward kernel.

    x_k, log_r, reg = local_step_autograd(eta1, eta2_diag, eta1_phi2, L_raw, pi_raw, theta_rec, S, seed=...)
    loss = -(decoder_loglike(x_k gathered at z) - reg);  loss.backward()

Latent dimension <= 16 and K <= 256 in the backward (the reference's training shapes are D = 2, 6; K = 10).
"""
import torch

from . import core


class _LocalStepFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eta1, eta2_diag, eta1_phi2, L_raw, pi_raw, theta_rec, S, den_mode, seed, noise):
        eta1, eta2_diag = eta1.contiguous(), eta2_diag.contiguous()
        eta1_phi2, L_raw, pi_raw = eta1_phi2.contiguous(), L_raw.contiguous(), pi_raw.contiguous()
        phi_rec = core.phi_prepare(eta1_phi2, L_raw, pi_raw)
        out = core.local_step(eta1, eta2_diag, phi_rec, theta_rec, S, den_mode=den_mode, noise=noise, seed=seed,
                              want_x_sample=False, want_z=False, materialize_x_k=True)
        ctx.save_for_backward(eta1, eta2_diag, eta1_phi2, L_raw, pi_raw, phi_rec, theta_rec, out['log_r'])
        ctx.noise, ctx.S, ctx.den_mode, ctx.seed = noise, S, den_mode, seed
        reg = out['elbo_acc'][2].to(eta1.dtype)
        ctx.mark_non_differentiable(out['elbo_acc'])
        return out['x_k_samples'], out['log_r'], reg, out['elbo_acc']

    @staticmethod
    def backward(ctx, gx, glr, greg, _gacc):
        eta1, eta2_diag, eta1_phi2, L_raw, pi_raw, phi_rec, theta_rec, log_r = ctx.saved_tensors
        N, D = eta1.shape
        K = phi_rec.shape[0]
        gx = torch.zeros(N, K, ctx.S, D, dtype=eta1.dtype, device=eta1.device) if gx is None else gx.contiguous()
        glr = torch.zeros(N, K, dtype=eta1.dtype, device=eta1.device) if glr is None else glr.contiguous()
        greg = 0.0 if greg is None else float(greg)     # one scalar sync; the loss weight is 1 or -1 in practice
        g = core.local_step_backward(eta1, eta2_diag, eta1_phi2, L_raw, pi_raw, phi_rec, theta_rec, ctx.S, log_r, gx,
                                     glr, greg, den_mode=ctx.den_mode, noise=ctx.noise, seed=ctx.seed)
        return g[0], g[1], g[2], g[3], g[4], None, None, None, None, None


def local_step_autograd(eta1, eta2_diag, eta1_phi2, L_raw, pi_raw, theta_rec, S, den_mode=core.DEN_GAUSS, seed=0,
                        noise=None):
    """-> (x_k_samples[N,K,S,D], log_r[N,K], regulariser (0-d), elbo_acc[4] double, non-differentiable).
    theta_rec comes from core.theta_prepare_* (a constant of the graph, as in the reference)."""
    return _LocalStepFn.apply(eta1, eta2_diag, eta1_phi2, L_raw, pi_raw, theta_rec, int(S), int(den_mode), int(seed),
                              noise)
