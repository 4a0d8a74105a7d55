"""Drop-in surface of the reference's models/svae.py on B200.

Same function names, argument order and return arity as the reference; `torch.Tensor` (CUDA, float32 or float64)
replaces `tf.Tensor`, eager execution replaces graph building, `name=` is accepted and ignored.  Keyword-only
additions: `noise=` / `u=` inject the random draws (TF's streams are not reproducible), `materialize=`
controls whether the O(N*K*S*D) / O(N*K*D*D) tensors are produced.

Arithmetic runs in libvmp_svae.so: e_step / subsample_x / compute_elbo(_smm) -> local_step.cu,
m_step -> suffstats.cu, unpack_* -> prepare.cu.  The fused, allocation-free version of the whole step (what
bench.py times) is `vmp_for_svae_b200.step.SVAEStep`.
"""
import math

import torch

from .. import core
from ..distributions import dirichlet, niw


# ------------------------------------------------------------------------------------------------ lazy outputs
class PhiTilde(object):
    """phi_tilde = (eta1_tilde[N,K,D,1], eta2_tilde[N,K,D,D]) of svae.py:39-42, kept as its two factors
    (encoder potentials + recognition-GMM records) and materialised only when unpacked / indexed."""

    def __init__(self, phi_enc, eta1_phi2, eta2_phi2, phi_rec):
        self.phi_enc = phi_enc
        self.eta1_phi2, self.eta2_phi2 = eta1_phi2, eta2_phi2
        self.phi_rec = phi_rec
        self._dense = None

    def dense(self):
        if self._dense is None:
            eta1_phi1, eta2_diag = self.phi_enc
            eta1 = (eta1_phi1.unsqueeze(1) + self.eta1_phi2.unsqueeze(0)).unsqueeze(-1)
            eta2 = torch.diag_embed(eta2_diag).unsqueeze(1) + self.eta2_phi2.unsqueeze(0)
            self._dense = (eta1, eta2)
        return self._dense

    def __iter__(self):
        return iter(self.dense())

    def __getitem__(self, i):
        return self.dense()[i]

    def __len__(self):
        return 2


class _LazyDbg(object):
    """dbg = (w_eta1[N,K,D], w_eta2[N,K,D,D]) of svae.py:73-86, computed on demand (debug output only):
    w_eta2 = -1/2 (Sigma1+Sigma2)^-1 = -1/2 (P1 - P1 P~^-1 P1), w_eta1 = (Sigma1+Sigma2)^-1 mu2."""

    def __init__(self, phi_tilde):
        self._pt = phi_tilde
        self._v = None

    def _get(self):
        if self._v is None:
            eta1_phi1, eta2_diag = self._pt.phi_enc
            _, eta2_tilde = self._pt.dense()
            p1 = -2.0 * eta2_diag                                            # N,D
            Pt_inv, _ = core.spd_inverse(-2.0 * eta2_tilde, want_logdet=False)   # N,K,D,D
            prec = torch.diag_embed(p1).unsqueeze(1) - p1[:, None, :, None] * Pt_inv * p1[:, None, None, :]
            D = eta1_phi1.shape[1]
            rec = self._pt.phi_rec
            mu2 = rec[:, D * D:D * D + D]
            self._v = (torch.einsum('nkij,kj->nki', prec, mu2), -0.5 * prec)
        return self._v

    def __iter__(self):
        return iter(self._get())

    def __getitem__(self, i):
        return self._get()[i]


def _zero_theta_rec(K, D, like):
    return torch.zeros(K, core._lib.record_lens(D)[1], dtype=like.dtype, device=like.device)


# ------------------------------------------------------------------------------------------------ e-step
def unpack_recognition_gmm(phi_gmm, name='unpack_phi2'):
    """svae.py:342-358 -> (eta1, eta2 = -1/2 L L^T, pi = softmax(pi_raw))."""
    eta1, L_k_raw, pi_k_raw = phi_gmm
    rec = core.phi_prepare(eta1, L_k_raw, pi_k_raw)
    K, D = eta1.shape
    P2 = rec[:, :D * D].reshape(K, D, D)
    return eta1, -0.5 * P2, torch.exp(rec[:, D * D + 2 * D])


def unpack_smm(theta_smm, name='unpack_theta_smm'):
    """svae.py:361-373 -> (mu, Sigma = L L^T)."""
    mu, L_k_raw = theta_smm
    K, D = mu.shape
    rec = core.phi_prepare(torch.zeros_like(mu), L_k_raw, torch.zeros(K, dtype=mu.dtype, device=mu.device))
    return mu, rec[:, :D * D].reshape(K, D, D).clone()


class NonPDFlag(object):
    """Device-side count of non-positive Cholesky pivots of an e_step call (the TF graph raises InvalidArgumentError
    for these at sess.run time).  Reading it is the only host synchronisation: `e_step` itself never syncs, so the
    surface can be captured in a CUDA graph; `raise_if_set()` (or e_step(..., check=True)) reproduces the raise."""

    def __init__(self, acc):
        self.acc = acc

    def count(self):
        return float(self.acc[3])

    def raise_if_set(self):
        if self.count() != 0.0:
            raise RuntimeError('Cholesky decomposition was not successful. The input might not be valid.')


def e_step(phi_enc, phi_gmm, nb_samples, seed=0, name="e_step", *, noise=None, u=None, materialize=True, check=False):
    """svae.py:14-47.

    Returns (x_k_samples[N,K,S,D], log_z_given_y_phi[N,K], phi_tilde, dbg).  `phi_tilde` and `dbg` are lazy
    (unpacking them materialises the [N,K,D,D] tensors); with materialize=False x_k_samples is None.
    noise: eps[N,K,D,S] (the raw_noise of svae.py:113-114); default: in-kernel Philox keyed by `seed`.
    u: Gumbel uniforms [N,K] of the fused categorical draw (only used by the training graph's forward).
    No host synchronisation: a non-PD system is reported through `phi_tilde.non_pd` (NonPDFlag); check=True reads it
    here and raises like the reference's sess.run does."""
    eta1_phi1, eta2_phi1_diag = phi_enc
    N, D = eta1_phi1.shape
    assert tuple(eta2_phi1_diag.shape) == (N, D)
    eta1_phi2, L_k_raw, pi_k_raw = phi_gmm
    K, L2 = eta1_phi2.shape
    assert L2 == D
    assert tuple(L_k_raw.shape) == (K, D, D)
    leaves = (eta1_phi1, eta2_phi1_diag, eta1_phi2, L_k_raw, pi_k_raw)
    if torch.is_grad_enabled() and any(t.requires_grad for t in leaves) and materialize:
        # training graph: the same forward kernel behind torch.autograd (reverse pass: csrc/local_step_bwd.cu)
        from ..autograd import local_step_autograd
        x_k, log_r, _, acc = local_step_autograd(*leaves, _zero_theta_rec(K, D, eta1_phi1), int(nb_samples), seed=seed,
                                                 noise=noise, u=u)
        phi_rec = core.phi_prepare(eta1_phi2.detach(), L_k_raw.detach(), pi_k_raw.detach())
        out = dict(x_k_samples=x_k, log_r=log_r, elbo_acc=acc)
    else:
        leaves = None
        phi_rec = core.phi_prepare(eta1_phi2, L_k_raw, pi_k_raw)
        out = core.local_step(eta1_phi1, eta2_phi1_diag, phi_rec, _zero_theta_rec(K, D, eta1_phi1), int(nb_samples),
                              noise=noise, u=u, seed=seed, want_x_sample=False, want_z=False,
                              materialize_x_k=bool(materialize))
    eta2_phi2 = -0.5 * phi_rec[:, :D * D].reshape(K, D, D)
    phi_tilde = PhiTilde((eta1_phi1.detach(), eta2_phi1_diag.detach()), eta1_phi2.detach(), eta2_phi2, phi_rec)
    phi_tilde.non_pd = NonPDFlag(out['elbo_acc'])
    if check:
        phi_tilde.non_pd.raise_if_set()
    phi_tilde.seed, phi_tilde.noise = seed, noise
    # autograd bookkeeping: compute_elbo re-enters the fused step with theta to get a differentiable regulariser
    phi_tilde.leaves, phi_tilde.x_k, phi_tilde.S = leaves, out['x_k_samples'], int(nb_samples)
    return out['x_k_samples'], out['log_r'], phi_tilde, _LazyDbg(phi_tilde)


def compute_log_z_given_y(eta1_phi1, eta2_phi1, eta1_phi2, eta2_phi2, pi_phi2, name='log_q_z_given_y_phi'):
    """svae.py:50-92 with the reference's argument layout (dense eta2_phi1[N,D,D] that must be diagonal, unpacked
    eta2_phi2 = -1/2 P2 and mixture weights pi).  Returns (log q(z|y,phi)[N,K], (w_eta1, w_eta2))."""
    N, L = eta1_phi1.shape
    assert tuple(eta2_phi1.shape) == (N, L, L)
    K, L2 = eta1_phi2.shape
    assert L2 == L
    assert tuple(eta2_phi2.shape) == (K, L, L)
    eta2_diag = torch.diagonal(eta2_phi1, dim1=-2, dim2=-1).contiguous()
    # rebuild the record from the unpacked parameters: P2 = -2 eta2, mu2 = P2^-1 eta1, logdet P2, log pi
    P2 = -2.0 * eta2_phi2
    P2_inv, logdet = core.spd_inverse(P2)
    mu2 = torch.einsum('kij,kj->ki', P2_inv, eta1_phi2)
    plen = core._lib.record_lens(L)[0]
    rec = torch.zeros(K, plen, dtype=eta1_phi1.dtype, device=eta1_phi1.device)
    rec[:, :L * L] = P2.reshape(K, L * L)
    rec[:, L * L:L * L + L] = mu2
    rec[:, L * L + L:L * L + 2 * L] = eta1_phi2
    rec[:, L * L + 2 * L] = torch.log(pi_phi2)
    rec[:, L * L + 2 * L + 1] = logdet
    out = core.local_step(eta1_phi1.contiguous(), eta2_diag, rec, _zero_theta_rec(K, L, eta1_phi1), 1, seed=0,
                          want_x_sample=False, want_z=False)
    pt = PhiTilde((eta1_phi1, eta2_diag), eta1_phi2, eta2_phi2, rec)
    return out['log_r'], _LazyDbg(pt)


def sample_x_per_comp(eta1, eta2, nb_samples, seed=0, *, noise=None):
    """svae.py:95-119 for dense eta1[N,K,D,1], eta2[N,K,D,D] (general API form; the fused path never builds these):
    x = P^-1 eta1 + L^-T eps, L = chol(P), P = -2 eta2 -> [N,K,S,D]  (vmp_gaussian_sample_nat, csrc/prepare.cu)."""
    N, K, _, D = eta2.shape
    if noise is None:
        noise, _ = core.fill_noise(N, K, D, int(nb_samples), seed, eta2.dtype, eta2.device, want_u=False)
    x, _ = core.gaussian_sample_nat(eta1, eta2.contiguous(), noise)
    return x


def subsample_x(x_k_samples, log_q_z_given_y, seed=0, *, cdf_u=None, gumbel_u=None, u=None):
    """svae.py:122-151 : x_samples[n,s] = x_k_samples[n, z_ns, s], z_ns ~ Cat(softmax(log q)) (tf.multinomial).
    Default / gumbel_u[N,S,K] (or [N,K] for S = 1): Gumbel-max, z = argmax_k(log q - log(-log u)) (TF's GPU kernel; for
    s = 0 and the same `seed` this is the draw the fused step makes).  cdf_u[N,S]: inverse-CDF search in double (TF's CPU
    kernel).  `u=` is the old name of cdf_u.  Stand-alone form of the gather for the reference's call order; the fused
    step selects inside the kernel."""
    N, K, S, L = x_k_samples.shape
    dev = x_k_samples.device
    if cdf_u is None:
        cdf_u = u
    if gumbel_u is not None and gumbel_u.dim() == 2:
        assert tuple(gumbel_u.shape) == (N, K) and S == 1, 'gumbel_u[N,K] needs one sample per component'
        gumbel_u = gumbel_u.unsqueeze(1)
    u = cdf_u
    if u is not None:
        assert tuple(u.shape) == (N, S), 'cdf_u must be [N,S] (inverse-CDF uniforms); Gumbel uniforms go in gumbel_u'

        lg = log_q_z_given_y.to(torch.float64)
        cdf = torch.cumsum(torch.exp(lg - lg.max(dim=1, keepdim=True).values), dim=1)
        z = torch.searchsorted(cdf, u.to(torch.float64) * cdf[:, -1:], right=True).clamp_(max=K - 1)
    else:
        if gumbel_u is None:
            gumbel_u = torch.stack([core.fill_noise(N, K, 1, 1, seed + 7919 * s, x_k_samples.dtype, dev,
                                                    want_noise=False)[1] for s in range(S)], dim=1)
        z = torch.argmax(log_q_z_given_y.unsqueeze(1) - torch.log(-torch.log(gumbel_u)), dim=2)
    n_idx = torch.arange(N, device=dev).reshape(-1, 1).expand(N, S)
    s_idx = torch.arange(S, device=dev).reshape(1, -1).expand(N, S)
    return x_k_samples[n_idx, z, s_idx]


# ------------------------------------------------------------------------------------------------ m-step / CVI
def m_step(gmm_prior, x_samples, r_nk):
    """svae.py:154-176 -> theta_star = [alpha, A, b, beta, v_hat] (natural parameters).
    One responsibility-weighted reduction (suffstats.cu) + the additive natural-parameter form of Bishop's M-step."""
    stats = core.suffstats(x_samples, r_nk, r_is_log=False)
    scratch = [t.clone() for t in gmm_prior]
    return core.ng_update(stats, 0.0, gmm_prior, scratch, want_star=True)


def m_step_smm(smm_prior, r_nk):
    """svae.py:179-196 -> alpha_star (natural parameter)."""
    alpha0 = smm_prior[0]
    N, K = r_nk.shape
    stats = core.suffstats(torch.zeros(N, 1, dtype=r_nk.dtype, device=r_nk.device), r_nk, r_is_log=False)
    scratch = [alpha0.clone()]
    return core.ng_update(stats, 0.0, [alpha0], scratch, only_alpha=True, want_star=True)[0]


def update_gmm_params(current_gmm_params, gmm_params_star, step_size, name='cvi_update_theta'):
    """svae.py:376-403 : in-place convex combination theta <- (1-rho) theta + rho theta* (K-sized axpy).
    The fused step applies this inside vmp_ng_update; this stand-alone form serves the reference's call order."""
    rho = float(step_size)
    for cur, star in zip(current_gmm_params, gmm_params_star):
        cur.mul_(1.0 - rho).add_(star, alpha=rho)
    return current_gmm_params


# ------------------------------------------------------------------------------------------------ ELBO
def _neg_reconstruction_error(y, reconstructions, r_nk, decoder_type):
    means, out_2 = reconstructions
    N, K, S, Dobs = out_2.shape
    if decoder_type not in ('standard', 'bernoulli'):
        raise NotImplementedError
    if torch.is_grad_enabled() and (out_2.requires_grad or r_nk.requires_grad or
                                    (means is not None and means.requires_grad)):
        from ..autograd import decoder_loglike_autograd
        return decoder_loglike_autograd(y, reconstructions, r_nk, decoder_type)
    if decoder_type == 'standard':
        acc = core.decoder_loglike(y, means, out_2, r_nk, 0)               # vae.py:226-248
        return (-0.5 * acc / S - N * Dobs / 2.0 * math.log(2.0 * math.pi)).to(out_2.dtype)[0]
    if decoder_type == 'bernoulli':
        acc = core.decoder_loglike(y, None, out_2, r_nk, 1)                # vae.py:175-198
        return (acc / S).to(out_2.dtype)[0]
    raise NotImplementedError


def _regulariser(theta_rec, den_mode, phi_tilde, x_k_samps, log_z_given_y_phi):
    N, K, S, L = x_k_samps.shape
    dt = x_k_samps.dtype
    if isinstance(phi_tilde, PhiTilde) and getattr(phi_tilde, 'leaves', None) is not None and torch.is_grad_enabled():
        # differentiable regulariser: the fused step again, now with theta (total derivative w.r.t. the leaves of
        # e_step; the samples are regenerated from the same noise, so x_k_samps must be e_step's own output)
        if x_k_samps is not phi_tilde.x_k:
            raise ValueError('under autograd compute_elbo needs the x_k_samples returned by e_step/inference')
        from ..autograd import local_step_autograd
        _, _, reg, acc = local_step_autograd(*phi_tilde.leaves, theta_rec, S, den_mode=den_mode, seed=phi_tilde.seed,
                                             noise=phi_tilde.noise)
        return reg, acc[0].to(dt), acc[1].to(dt)
    if isinstance(phi_tilde, PhiTilde):
        eta1, eta2_diag = phi_tilde.phi_enc
        out = core.local_step(eta1, eta2_diag, phi_tilde.phi_rec, theta_rec, S, den_mode=den_mode,
                              x_in=x_k_samps.contiguous(), want_x_sample=False, want_z=False)
        acc = out['elbo_acc']
        return acc[2].to(dt), acc[0].to(dt), acc[1].to(dt)
    # dense phi_tilde supplied by the caller: general-form kernels (distributions.gaussian)
    from ..distributions import gaussian
    eta1_t, eta2_t = phi_tilde
    eta1_t = eta1_t.reshape(N, K, L)
    r_nk = torch.exp(log_z_given_y_phi)
    log_num = gaussian.log_probability_nat_per_samp(x_k_samps, eta1_t, eta2_t) + log_z_given_y_phi.unsqueeze(2)
    D = L
    W = theta_rec[:, :D * D].reshape(K, D, D)
    m = theta_rec[:, D * D:D * D + D]
    q = torch.einsum('kij,nksj->nksi', W, x_k_samps - m.unsqueeze(0).unsqueeze(2))
    maha = (q * q).sum(-1)
    cden, nu = theta_rec[:, D * D + D], theta_rec[:, D * D + D + 1]
    if den_mode == core.DEN_GAUSS:
        log_den = cden.reshape(1, K, 1) - 0.5 * maha
    else:
        log_den = cden.reshape(1, K, 1) - 0.5 * (nu.reshape(1, K, 1) + D) * torch.log1p(maha / nu.reshape(1, K, 1))
    reg = (r_nk.unsqueeze(2) * (log_num - log_den)).sum(1).sum(0).mean()
    return reg, (r_nk * log_num.mean(-1)).sum(), (r_nk * log_den.mean(-1)).sum()


def compute_elbo(y, reconstructions, theta, phi_tilde, x_k_samps, log_z_given_y_phi, decoder_type):
    """svae.py:199-262 -> (elbo, (neg_rec, sum r*mean_s log num, sum r*mean_s log den, regulariser))."""
    theta_rec = core.theta_prepare_gauss(theta)
    r_nk = torch.exp(log_z_given_y_phi)
    neg_rec = _neg_reconstruction_error(y, reconstructions, r_nk, decoder_type)
    reg, num, den = _regulariser(theta_rec, core.DEN_GAUSS, phi_tilde, x_k_samps, log_z_given_y_phi)
    return neg_rec - reg, (neg_rec, num, den, reg)


def compute_elbo_smm(y, reconstructions, theta, phi_tilde, x_k_samps, log_z_given_y_phi, decoder_type):
    """svae.py:265-322 ; theta = (alpha_nat, mu_k, L_k_raw, DoF) (experiments.py:174)."""
    if torch.is_grad_enabled() and (theta[1].requires_grad or theta[2].requires_grad):
        from ..autograd import student_theta_record             # mu_k, L_k are trained by gradient (experiments.py:154-174)
        theta_rec = student_theta_record(theta[0], theta[1], theta[2], theta[3])
    else:
        theta_rec = core.theta_prepare_student(theta)
    r_nk = torch.exp(log_z_given_y_phi)
    neg_rec = _neg_reconstruction_error(y, reconstructions, r_nk, decoder_type)
    reg, num, den = _regulariser(theta_rec, core.DEN_STUDENT, phi_tilde, x_k_samps, log_z_given_y_phi)
    return neg_rec - reg, (neg_rec, num, den, reg)


# ------------------------------------------------------------------------------------------------ prediction
def predict(y, phi_gmm, encoder, decoder, seed=0):
    """svae.py:406-430 with torch callables for the encoder / decoder (they stay ordinary PyTorch).
    Returns (y_mean, argmax_k log r)."""
    phi_enc = encoder(y)
    eta1, eta2_diag = phi_enc
    K, D = phi_gmm[0].shape
    phi_rec = core.phi_prepare(*phi_gmm)
    out = core.local_step(eta1.contiguous(), eta2_diag.contiguous(), phi_rec, _zero_theta_rec(K, D, eta1), 1, seed=seed)
    y_mean, _ = decoder(out['x_sample'])
    return y_mean, torch.argmax(out['log_r'], dim=1)


def inference(y, phi_gmm, encoder, decoder, nb_samples=10, stddev_init_nn=0.01, seed=0, name='inference',
              param_device=None, *, noise=None, cdf_u=None, gumbel_u=None, u=None):
    """svae.py:499-516 -> (y_reconstruction, x_given_y_phi, x_k_samples, x_samples, log_z_given_y_phi, phi_gmm,
    phi_tilde).  encoder / decoder are torch callables (the reference builds them from layer specs).
    The categorical draw of subsample_x: gumbel_u[N,S,K] (Gumbel-max) or cdf_u[N,S] (inverse CDF; `u=` is its old name)."""
    x_given_y_phi = encoder(y)
    x_given_y_phi = (x_given_y_phi[0].contiguous(), x_given_y_phi[1].contiguous())
    x_k_samples, log_z_given_y_phi, phi_tilde, _ = e_step(x_given_y_phi, phi_gmm, nb_samples, seed=seed, noise=noise)
    y_reconstruction = decoder(x_k_samples)
    x_samples = subsample_x(x_k_samples, log_z_given_y_phi, seed, cdf_u=cdf_u if cdf_u is not None else u,
                            gumbel_u=gumbel_u)[:, 0, :]
    return y_reconstruction, x_given_y_phi, x_k_samples, x_samples, log_z_given_y_phi, phi_gmm, phi_tilde


# ------------------------------------------------------------------------------------------------ initialisation
def init_mm_params(nb_components, latent_dims, alpha_scale=.1, beta_scale=1e-5, v_init=10., m_scale=1., C_scale=10.,
                   seed=0, as_variables=True, trainable=False, device='cuda', name='gmm', dtype=torch.float32,
                   uniform=None):
    """svae.py:433-458 -> (alpha, A, b, beta, v_hat) natural parameters.  `uniform` injects the U[0,1) draws."""
    K, D = nb_components, latent_dims
    alpha_init = alpha_scale * torch.ones(K, dtype=dtype, device=device)
    beta_init = beta_scale * torch.ones(K, dtype=dtype, device=device)
    v = torch.full((K,), float(D + v_init), dtype=dtype, device=device)
    if uniform is None:
        g = torch.Generator(device='cpu').manual_seed(int(seed))
        uniform = torch.rand(K, D, generator=g, dtype=torch.float64)
    means_init = m_scale * (-1.0 + 2.0 * uniform.to(device=device, dtype=dtype))
    covariance_init = C_scale * torch.eye(D, dtype=dtype, device=device).unsqueeze(0).repeat(K, 1, 1)
    A, b, beta, v_hat = niw.standard_to_natural(beta_init, means_init, covariance_init, v)
    alpha = dirichlet.standard_to_natural(alpha_init)
    return [alpha.contiguous(), A.contiguous(), b.contiguous(), beta.contiguous(), v_hat.contiguous()]


def init_mm(nb_components, latent_dims, seed=0, param_device='cuda', name='init_mm', theta_as_variable=True,
            dtype=torch.float32, uniform=None):
    """svae.py:461-471 -> (theta_prior, theta)."""
    theta_prior = init_mm_params(nb_components, latent_dims, alpha_scale=0.05 / nb_components, beta_scale=0.5,
                                 m_scale=0, C_scale=latent_dims + 0.5, v_init=latent_dims + 0.5, seed=seed,
                                 device=param_device, dtype=dtype, uniform=uniform)
    theta = init_mm_params(nb_components, latent_dims, alpha_scale=1., beta_scale=1., m_scale=5.,
                           C_scale=2 * latent_dims, v_init=latent_dims + 1., seed=seed, device=param_device,
                           dtype=dtype, uniform=uniform)
    return theta_prior, theta


def make_loc_scale_variables(theta, param_device=None, name='copy_m_v'):
    """svae.py:474-485 : mu_k = E[mu], L_k = chol(E[Sigma]) (K-sized initialisation)."""
    theta_copied = niw.natural_to_standard(theta[1].clone(), theta[2].clone(), theta[3].clone(), theta[4].clone())
    mu_k_init, sigma_k = niw.expected_values(theta_copied)
    return mu_k_init.contiguous(), torch.linalg.cholesky(sigma_k).contiguous()


def init_recognition_params(theta, nb_components, seed=0, param_device=None, var_scope='phi_gmm', normal=None):
    """svae.py:488-496 -> (mu_k, L_k, pi_k); note pi_k is already a softmax and is softmaxed again on use
    (svae.py:491 vs 356) — reproduced verbatim."""
    dtype, device = theta[1].dtype, theta[1].device
    if normal is None:
        g = torch.Generator(device='cpu').manual_seed(int(seed))
        normal = torch.randn(nb_components, generator=g, dtype=torch.float64)
    pi_k_init = torch.softmax(normal.to(device=device, dtype=dtype), dim=-1)
    mu_k, L_k = make_loc_scale_variables(theta, param_device)
    return mu_k, L_k, pi_k_init.contiguous()
