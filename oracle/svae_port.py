"""ORACLE (test infrastructure, not product code): literal CPU restatement of the reference's
SVAE local/global step (models/svae.py) — the same op sequence as the TF 1.3 graph:
materialised [N,K,D,D] tensors, batched LU solves, batched Choleskys, einsums.
torch-CPU, dtype-generic.  Used (1) as the parity checker in tests/ and smoke(),
(2) timed in fp32 with all host threads as bench.py's cpu_baseline / `--impl reference`
("port": restatement of the TF 1.3 graph, not TensorFlow itself).

Randomness is injected (TF's Philox streams are not reproducible):
  noise : eps[N,K,D,S]  = the `raw_noise` of svae.py:113-114
  u     : uniforms[N,S] = the draws inside tf.multinomial (svae.py:142); z is the CPU
          kernel's inverse-CDF pick: upper_bound(cumsum(exp(logit-max)), u*total), in double.

Pinned against the reference's own source executed over oracle/tf_shim
(tests/golden/*.npz); TensorFlow itself cannot run here, see DESIGN.md.
"""
import math

import torch

from . import dists, mixtures


def softplus(x):
    return torch.nn.functional.softplus(x, beta=1.0, threshold=1e6)


def unpack_recognition_gmm(phi_gmm):
    """svae.py:342-358 : (eta1, L_raw, pi_raw) -> (eta1, eta2=-1/2 L L^T, softmax(pi_raw)),
    L = tril(L_raw) with softplus on the diagonal."""
    eta1, L_raw, pi_raw = phi_gmm
    L = torch.tril(L_raw)
    L = L - torch.diag_embed(torch.diagonal(L, dim1=-2, dim2=-1)) \
        + torch.diag_embed(softplus(torch.diagonal(L, dim1=-2, dim2=-1)))
    P = L @ L.transpose(-1, -2)
    return eta1, -0.5 * P, torch.softmax(pi_raw, dim=-1)


def unpack_smm(theta_smm):
    """svae.py:361-373 : (mu, L_raw) -> (mu, Sigma = L L^T)."""
    mu, L_raw = theta_smm
    L = torch.tril(L_raw)
    L = L - torch.diag_embed(torch.diagonal(L, dim1=-2, dim2=-1)) \
        + torch.diag_embed(softplus(torch.diagonal(L, dim1=-2, dim2=-1)))
    return mu, L @ L.transpose(-1, -2)


def compute_log_z_given_y(eta1_phi1, eta2_phi1, eta1_phi2, eta2_phi2, pi_phi2):
    """svae.py:50-92 : two batched LU solves of eta2_tilde[N,K,D,D], symmetrisation,
    diagonal inverse for mu_phi1, then gaussian.log_probability_nat."""
    N, L = eta1_phi1.shape
    assert tuple(eta2_phi1.shape) == (N, L, L)
    K, L2 = eta1_phi2.shape
    assert L2 == L
    assert tuple(eta2_phi2.shape) == (K, L, L)
    eta2_tilde = eta2_phi1.unsqueeze(1) + eta2_phi2.unsqueeze(0)
    solved = torch.linalg.solve(eta2_tilde, eta2_phi2.unsqueeze(0).expand(N, K, L, L))
    w_eta2 = torch.einsum('nju,nkui->nkij', eta2_phi1, solved)
    w_eta2 = (w_eta2 + w_eta2.transpose(-1, -2)) / 2.0
    rhs = eta1_phi2.unsqueeze(0).unsqueeze(-1).expand(N, K, L, 1)
    w_eta1 = torch.einsum('nuj,nkuv->nkj', eta2_phi1, torch.linalg.solve(eta2_tilde, rhs))
    mu_phi1, _ = dists.gaussian_natural_to_standard(eta1_phi1, eta2_phi1)
    return dists.gaussian_log_probability_nat(mu_phi1, w_eta1, w_eta2, pi_phi2), (w_eta1, w_eta2)


def sample_x_per_comp(eta1, eta2, noise):
    """svae.py:95-119 with the raw noise injected: eta1[N,K,D,1], eta2[N,K,D,D], noise[N,K,D,S]
    -> x[N,K,S,D] = transpose(solve(P, eta1) + solve(L^T, noise)), L = chol(P), P = -2 eta2."""
    inv_sigma = -2.0 * eta2
    Lc = torch.linalg.cholesky(inv_sigma)
    nz = torch.linalg.solve(Lc.transpose(-1, -2), noise)
    return (torch.linalg.solve(inv_sigma, eta1) + nz).permute(0, 1, 3, 2)


def multinomial_inverse_cdf(logits, u):
    """tf.multinomial's CPU kernel (multinomial_op.cc) with injected uniforms u[N,S]."""
    lg = logits.to(torch.float64)
    mx = lg.max(dim=1, keepdim=True).values
    cdf = torch.cumsum(torch.exp(lg - mx), dim=1)
    target = u.to(torch.float64) * cdf[:, -1:]
    z = torch.searchsorted(cdf.contiguous(), target.contiguous(), right=True)
    return z.clamp_(max=logits.shape[1] - 1)


def multinomial_gumbel_max(logits, gumbel_u):
    """tf.multinomial's GPU kernel (multinomial_op_gpu.cu.cc) with injected uniforms gumbel_u[N,S,K]:
    z_ns = argmax_k(logit_nk - log(-log(u_nsk))) (first maximum on ties)."""
    g = -torch.log(-torch.log(gumbel_u.to(logits.dtype)))
    return torch.argmax(logits.unsqueeze(1) + g, dim=2)


def subsample_x(x_k_samples, log_q_z_given_y, u=None, gumbel_u=None):
    """svae.py:122-151 : pick x_k_samples[n, z_ns, s], z_ns ~ Cat(softmax(log q)) -> [N,S,L].
    The draw is tf.multinomial: its CPU kernel is an inverse-CDF search (inject u[N,S]), its GPU kernel a
    Gumbel-max (inject gumbel_u[N,S,K]); both sample the same categorical."""
    N, K, S, L = x_k_samples.shape
    if gumbel_u is not None:
        z = multinomial_gumbel_max(log_q_z_given_y, gumbel_u)
    else:
        z = multinomial_inverse_cdf(log_q_z_given_y, u)                  # N,S
    n_idx = torch.arange(N).reshape(-1, 1).expand(N, S)
    s_idx = torch.arange(S).reshape(1, -1).expand(N, S)
    return x_k_samples[n_idx, z, s_idx], z


def e_step(phi_enc, phi_gmm, noise):
    """svae.py:14-47 -> (x_k_samples[N,K,S,D], log_z_given_y_phi[N,K], phi_tilde, dbg)."""
    eta1_phi1, eta2_phi1_diag = phi_enc
    eta2_phi1 = torch.diag_embed(eta2_phi1_diag)
    eta1_phi2, eta2_phi2, pi_phi2 = unpack_recognition_gmm(phi_gmm)
    log_z, dbg = compute_log_z_given_y(eta1_phi1, eta2_phi1, eta1_phi2, eta2_phi2, pi_phi2)
    eta1_tilde = (eta1_phi1.unsqueeze(1) + eta1_phi2.unsqueeze(0)).unsqueeze(-1)
    eta2_tilde = eta2_phi1.unsqueeze(1) + eta2_phi2.unsqueeze(0)
    x_k = sample_x_per_comp(eta1_tilde, eta2_tilde, noise)
    return x_k, log_z, (eta1_tilde, eta2_tilde), dbg


def m_step(gmm_prior, x_samples, r_nk):
    """svae.py:154-176 -> theta_star = [alpha, A, b, beta, v_hat] (natural parameters)."""
    beta_0, m_0, C_0, v_0 = dists.niw_natural_to_standard(*gmm_prior[1:])
    alpha_0 = dists.dirichlet_natural_to_standard(gmm_prior[0])
    alpha_k, beta_k, m_k, C_k, v_k, _, _ = mixtures.gmm_m_step(x_samples, r_nk, alpha_0, beta_0, m_0, C_0, v_0)
    A, b, beta, v_hat = dists.niw_standard_to_natural(beta_k, m_k, C_k, v_k)
    return [dists.dirichlet_standard_to_natural(alpha_k), A, b, beta, v_hat]


def m_step_smm(smm_prior, r_nk):
    """svae.py:179-196."""
    alpha_0 = dists.dirichlet_natural_to_standard(smm_prior[0])
    return dists.dirichlet_standard_to_natural(alpha_0 + r_nk.sum(0))


def update_gmm_params(current, star, step_size):
    """svae.py:376-403 : in-place convex combination theta <- (1-rho) theta + rho theta*."""
    for cur, st in zip(current, star):
        cur.copy_((1 - step_size) * cur + step_size * st)
    return current


def cvi_step_size(lrcvi, global_step, decay_rate, decay_steps=1000):
    """experiments.py:146 : tf.train.exponential_decay(lrcvi, step, 1000, decay_rate)."""
    return lrcvi * decay_rate ** (float(global_step) / decay_steps)


def expected_diagonal_gaussian_loglike(y, means, vars_, weights=None):
    """models/vae.py:201-250."""
    if weights is None:
        if means.dim() != 3:
            means, vars_ = means.unsqueeze(1), vars_.unsqueeze(1)
        M, S, L = means.shape
        sm = (((y.unsqueeze(1) - means) ** 2) / vars_).sum() + torch.log(vars_).sum()
    else:
        M, K, S, L = means.shape
        yy = y.unsqueeze(1).unsqueeze(1)
        sm = torch.einsum('nksd,nk->', (yy - means) ** 2 / vars_ + torch.log(vars_ + 1e-8), weights)
    return -0.5 * sm / S - M * L / 2.0 * math.log(2.0 * math.pi)


def expected_bernoulli_loglike(y_binary, logits, r_nk=None):
    """models/vae.py:175-198."""
    yb = y_binary.unsqueeze(1)
    if r_nk is not None:
        yb = yb.unsqueeze(1)
    pix = -torch.log(1.0 + torch.exp(-logits * yb))
    img = pix.sum(-1).mean(-1)
    if r_nk is not None:
        img = (r_nk * img).sum(1)
    return img.sum()


def _regulariser(log_num, log_den, r_nk):
    """svae.py:245-260 (shared by both ELBOs)."""
    reg = (r_nk.unsqueeze(2) * (log_num - log_den)).sum(1).sum(0).mean()
    return reg, (r_nk * log_num.mean(-1)).sum(), (r_nk * log_den.mean(-1)).sum()


def _neg_rec(y, reconstructions, r_nk, decoder_type):
    means, out_2 = reconstructions
    if decoder_type == 'standard':
        return expected_diagonal_gaussian_loglike(y, means, out_2, weights=r_nk)
    if decoder_type == 'bernoulli':
        return expected_bernoulli_loglike(y, out_2, r_nk=r_nk)
    raise NotImplementedError


def compute_elbo(y, reconstructions, theta, phi_tilde, x_k_samps, log_z_given_y_phi, decoder_type):
    """svae.py:199-262 -> (elbo, (neg_rec, sum r*mean_s num, sum r*mean_s den, regulariser))."""
    beta_k, m_k, C_k, v_k = dists.niw_natural_to_standard(*theta[1:])
    mu, sigma = dists.niw_expected_values((beta_k, m_k, C_k, v_k))
    eta1_theta, eta2_theta = dists.gaussian_standard_to_natural(mu, sigma)
    e_log_pi = dists.dirichlet_expected_log_pi(dists.dirichlet_natural_to_standard(theta[0]))
    r_nk = torch.exp(log_z_given_y_phi)
    neg_rec = _neg_rec(y, reconstructions, r_nk, decoder_type)
    eta1_t, eta2_t = phi_tilde
    N, K, L, _ = eta2_t.shape
    eta1_t = eta1_t.reshape(N, K, L)
    log_num = dists.gaussian_log_probability_nat_per_samp(x_k_samps, eta1_t, eta2_t) + log_z_given_y_phi.unsqueeze(2)
    log_den = dists.gaussian_log_probability_nat_per_samp(
        x_k_samps, eta1_theta.unsqueeze(0).expand(N, K, L), eta2_theta.unsqueeze(0).expand(N, K, L, L))
    log_den = log_den + e_log_pi.unsqueeze(0).unsqueeze(2)
    reg, num_s, den_s = _regulariser(log_num, log_den, r_nk)
    return neg_rec - reg, (neg_rec, num_s, den_s, reg)


def compute_elbo_smm(y, reconstructions, theta, phi_tilde, x_k_samps, log_z_given_y_phi, decoder_type):
    """svae.py:265-322 ; theta = (alpha_nat, mu_k, L_k_raw, dof)."""
    mu_theta, sigma_theta = unpack_smm(theta[1:3])
    e_log_pi = dists.dirichlet_expected_log_pi(dists.dirichlet_natural_to_standard(theta[0]))
    dof = theta[3]
    r_nk = torch.exp(log_z_given_y_phi)
    neg_rec = _neg_rec(y, reconstructions, r_nk, decoder_type)
    eta1_t, eta2_t = phi_tilde
    N, K, L, _ = eta2_t.shape
    eta1_t = eta1_t.reshape(N, K, L)
    log_num = dists.gaussian_log_probability_nat_per_samp(x_k_samps, eta1_t, eta2_t) + log_z_given_y_phi.unsqueeze(2)
    log_den = dists.student_t_log_probability_per_samp(x_k_samps, mu_theta, sigma_theta, dof)
    log_den = log_den + e_log_pi.unsqueeze(0).unsqueeze(2)
    reg, num_s, den_s = _regulariser(log_num, log_den, r_nk)
    return neg_rec - reg, (neg_rec, num_s, den_s, reg)


# ---------------------------------------------------------------------------- initialisation
def init_mm_params(nb_components, latent_dims, alpha_scale=.1, beta_scale=1e-5, v_init=10., m_scale=1.,
                   C_scale=10., uniform=None, dtype=torch.float64):
    """svae.py:433-458 ; `uniform` = the U[0,1) draws behind tf.random_uniform (K,D)."""
    K, D = nb_components, latent_dims
    alpha = alpha_scale * torch.ones(K, dtype=dtype)
    beta = beta_scale * torch.ones(K, dtype=dtype)
    v = torch.full((K,), float(D + v_init), dtype=dtype)
    if uniform is None:
        uniform = torch.full((K, D), 0.5, dtype=dtype)
    means = m_scale * (-1.0 + 2.0 * uniform.to(dtype))
    cov = C_scale * torch.eye(D, dtype=dtype).unsqueeze(0).repeat(K, 1, 1)
    A, b, beta, v_hat = dists.niw_standard_to_natural(beta, means, cov, v)
    return [dists.dirichlet_standard_to_natural(alpha), A, b, beta, v_hat]


def init_mm(nb_components, latent_dims, uniform=None, dtype=torch.float64):
    """svae.py:461-471 -> (theta_prior, theta)."""
    K, D = nb_components, latent_dims
    prior = init_mm_params(K, D, alpha_scale=0.05 / K, beta_scale=0.5, m_scale=0, C_scale=D + 0.5,
                           v_init=D + 0.5, uniform=uniform, dtype=dtype)
    theta = init_mm_params(K, D, alpha_scale=1., beta_scale=1., m_scale=5., C_scale=2 * D, v_init=D + 1.,
                           uniform=uniform, dtype=dtype)
    return prior, theta


def make_loc_scale_variables(theta):
    """svae.py:474-485 : mu_k = E[mu], L_k = chol(E[Sigma])."""
    mu, sigma = dists.niw_expected_values(dists.niw_natural_to_standard(theta[1], theta[2], theta[3], theta[4]))
    return mu, torch.linalg.cholesky(sigma)


def init_recognition_params(theta, nb_components, normal=None):
    """svae.py:488-496 ; `normal` = the N(0,1) draws behind tf.random_normal (K,)."""
    dtype = theta[1].dtype
    if normal is None:
        normal = torch.zeros(nb_components, dtype=dtype)
    pi = torch.softmax(normal.to(dtype), dim=-1)
    mu_k, L_k = make_loc_scale_variables(theta)
    return mu_k, L_k, pi


# ---------------------------------------------------------------------------- whole step
def svae_step(phi_enc, phi_gmm, theta, prior, noise, u, rho, gumbel_u=None):
    """One pass of the hot path in the order the reference graph runs it
    (experiments.py:208-260; ELBO reads theta BEFORE the update):
      e_step -> subsample_x[:,0,:] -> ELBO regulariser -> m_step -> update_gmm_params.
    u[N,S]: inverse-CDF uniforms, or gumbel_u[N,K] (sample 0 only): Gumbel-max uniforms.
    Returns dict(log_r, x_samples, z, reg, num, den, theta_new)."""
    x_k, log_r, phi_tilde, _ = e_step(phi_enc, phi_gmm, noise)
    if gumbel_u is not None:
        S = x_k.shape[2]
        gu = gumbel_u.unsqueeze(1).expand(-1, S, -1)
        xs, z = subsample_x(x_k, log_r, gumbel_u=gu)
    else:
        xs, z = subsample_x(x_k, log_r, u)
    x_samples = xs[:, 0, :]
    N, K = log_r.shape
    dt = log_r.dtype
    y0 = torch.zeros(N, 1, dtype=dt)
    rec = (torch.zeros(N, K, x_k.shape[2], 1, dtype=dt), torch.ones(N, K, x_k.shape[2], 1, dtype=dt))
    _, (_, num_s, den_s, reg) = compute_elbo(y0, rec, theta, phi_tilde, x_k, log_r, 'standard')
    star = m_step(prior, x_samples, torch.exp(log_r))
    theta_new = [t.clone() for t in theta]
    update_gmm_params(theta_new, star, rho)
    return dict(log_r=log_r, x_samples=x_samples, z=z[:, 0], reg=reg, num=num_s, den=den_s,
                theta_new=theta_new, x_k=x_k, theta_star=star)
