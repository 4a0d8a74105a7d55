"""Differentiable fused local step: the piece of the reference's training graph between the encoder outputs and the
decoder input / ELBO regulariser, with the hand-written reverse pass (csrc/local_step_bwd.cu) behind
`torch.autograd.Function`.

Reference: experiments.py:208-232 builds  elbo = E log p(y|x) - regulariser  from svae.inference (svae.py:325-339) and
calls opt.compute_gradients(-elbo) on phi_enc / phi_gmm / decoder weights; TF differentiates through e_step
(svae.py:39-100), the per-component samples (svae.py:103-123) and compute_elbo (svae.py:199-262) with theta behind
tf.stop_gradient (svae.py:211-214).  Here those three are ONE forward kernel and ONE backward kernel.

    x_k, log_r, reg = local_step_autograd(eta1, eta2_diag, eta1_phi2, L_raw, pi_raw, theta_rec, S, seed=...)
    loss = -(decoder_loglike(x_k gathered at z) - reg);  loss.backward()

Latent dimension <= 64 in the backward (thread-per-pair kernel up to D = 16, block-cooperative kernel above).
"""
import torch

from . import core


class _LocalStepFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eta1, eta2_diag, eta1_phi2, L_raw, pi_raw, theta_rec, S, den_mode, seed, noise, u):
        eta1, eta2_diag = eta1.contiguous(), eta2_diag.contiguous()
        eta1_phi2, L_raw, pi_raw = eta1_phi2.contiguous(), L_raw.contiguous(), pi_raw.contiguous()
        phi_rec = core.phi_prepare(eta1_phi2, L_raw, pi_raw)
        theta_rec = theta_rec.detach().contiguous()
        out = core.local_step(eta1, eta2_diag, phi_rec, theta_rec, S, den_mode=den_mode, noise=noise, u=u, seed=seed,
                              want_x_sample=True, want_z=True, materialize_x_k=True)
        ctx.save_for_backward(eta1, eta2_diag, eta1_phi2, L_raw, pi_raw, phi_rec, theta_rec, out['log_r'])
        ctx.noise, ctx.S, ctx.den_mode, ctx.seed = noise, S, den_mode, seed
        reg = out['elbo_acc'][2].to(eta1.dtype)
        ctx.mark_non_differentiable(out['elbo_acc'], out['x_sample'], out['z'])
        return out['x_k_samples'], out['log_r'], reg, out['elbo_acc'], out['x_sample'], out['z']

    @staticmethod
    def backward(ctx, gx, glr, greg, _gacc, _gxs, _gz):
        eta1, eta2_diag, eta1_phi2, L_raw, pi_raw, phi_rec, theta_rec, log_r = ctx.saved_tensors
        N, D = eta1.shape
        K = phi_rec.shape[0]
        gx = torch.zeros(N, K, ctx.S, D, dtype=eta1.dtype, device=eta1.device) if gx is None else gx.contiguous()
        glr = torch.zeros(N, K, dtype=eta1.dtype, device=eta1.device) if glr is None else glr.contiguous()
        greg = 0.0 if greg is None else greg            # stays on the device: no host synchronisation
        want_th = ctx.needs_input_grad[5]
        g = core.local_step_backward(eta1, eta2_diag, eta1_phi2, L_raw, pi_raw, phi_rec, theta_rec, ctx.S, log_r, gx,
                                     glr, greg, den_mode=ctx.den_mode, noise=ctx.noise, seed=ctx.seed,
                                     want_theta_rec_bar=want_th)
        return g[0], g[1], g[2], g[3], g[4], (g[5] if want_th else None), None, None, None, None, None


def local_step_autograd(eta1, eta2_diag, eta1_phi2, L_raw, pi_raw, theta_rec, S, den_mode=core.DEN_GAUSS, seed=0,
                        noise=None, u=None, full=False):
    """-> (x_k_samples[N,K,S,D], log_r[N,K], regulariser (0-d), elbo_acc[4] double, non-differentiable); with
    full=True also the selected sample x[n, z_n, 0] [N,D] and z[N] of the same pass (inputs of the CVI M-step).
    theta_rec comes from core.theta_prepare_* (a constant of the graph for the GMM prior, as in the reference) or from
    `student_theta_record` below (differentiable: the SMM variant trains mu_k, L_k by gradient)."""
    out = _LocalStepFn.apply(eta1, eta2_diag, eta1_phi2, L_raw, pi_raw, theta_rec, int(S), int(den_mode), int(seed),
                             noise, u)
    return out if full else out[:4]


def student_theta_record(alpha_nat, mu_k, L_k_raw, dof):
    """Differentiable theta record of the Student-t denominator (svae.py:265-306; svae.unpack_smm 361-373):
    Sigma_k = L L^T with L = tril(L_raw), softplus on the diagonal, so chol(Sigma_k) = L and W_k = L^-1.  K-sized torch
    ops (plumbing); E log pi and the degrees of freedom carry no gradient (svae.py:275-277).  Same layout and values as
    core.theta_prepare_student."""
    K, D = mu_k.shape
    L = torch.tril(L_k_raw, -1) + torch.diag_embed(torch.nn.functional.softplus(torch.diagonal(L_k_raw, dim1=-2, dim2=-1)))
    W = torch.linalg.solve_triangular(L, torch.eye(D, dtype=mu_k.dtype, device=mu_k.device).expand(K, D, D), upper=False)
    alpha = (alpha_nat + 1.0).detach()
    e_log_pi = torch.digamma(alpha) - torch.digamma(alpha.sum())
    nu = dof.detach()
    logdet_half = torch.log(torch.diagonal(L, dim1=-2, dim2=-1)).sum(-1)
    import math
    cden = e_log_pi + torch.lgamma((nu + D) / 2.0) - torch.lgamma(nu / 2.0) - 0.5 * D * torch.log(nu * math.pi) - logdet_half
    return torch.cat([W.reshape(K, D * D), mu_k, cden.unsqueeze(1), nu.unsqueeze(1), e_log_pi.unsqueeze(1),
                      (-2.0 * logdet_half).unsqueeze(1)], dim=1)


class _DecoderLoglikeFn(torch.autograd.Function):
    """neg_rec term of the ELBO (svae.py:220-223): vae.expected_diagonal_gaussian_loglike (mode 0, vae.py:226-248) or
    vae.expected_bernoulli_loglike (mode 1, vae.py:175-198) with responsibilities as weights; one streaming kernel each
    way over the decoder outputs [N,K,S,Dobs]."""

    @staticmethod
    def forward(ctx, y, means, out2, w, mode):
        import math
        N, K, S, Dobs = out2.shape
        y, out2, w = y.contiguous(), out2.contiguous(), w.contiguous()
        means = means.contiguous() if mode == 0 else None
        acc = core.decoder_loglike(y, means, out2, w, mode)
        ctx.save_for_backward(y, means, out2, w)
        ctx.mode = mode
        ctx.scale = -0.5 / S if mode == 0 else 1.0 / S
        val = acc[0] * ctx.scale - (N * Dobs / 2.0 * math.log(2.0 * math.pi) if mode == 0 else 0.0)
        return val.to(out2.dtype)

    @staticmethod
    def backward(ctx, g):
        y, means, out2, w = ctx.saved_tensors
        g_means, g_out2, g_w = core.decoder_loglike_backward(y, means, out2, w, ctx.mode, ctx.scale)
        return None, (g_means * g if g_means is not None else None), g_out2 * g, g_w * g, None


def decoder_loglike_autograd(y, reconstructions, r_nk, decoder_type):
    """Differentiable `neg_rec` of svae.compute_elbo: reconstructions = (means, out_2) [N,K,S,Dobs], r_nk [N,K]."""
    means, out2 = reconstructions
    mode = {'standard': 0, 'bernoulli': 1}[decoder_type]
    if mode == 1 and means is None:
        means = out2
    return _DecoderLoglikeFn.apply(y, means, out2, r_nk, mode)
