"""ORACLE (test infrastructure): the differentiable form of the fused local step and its hand-derived backward.

`forward` is the one-Cholesky formulation of svae.e_step + the ELBO regulariser written with
differentiable torch ops, so `torch.autograd` gives reference gradients — this is what the reference obtains from
`opt.compute_gradients(-elbo)` (experiments.py:232) with theta behind `tf.stop_gradient` (svae.py:211-214).
`backward_closed_form` restates the reverse pass as explicit per-pair formulas; it is the specification the CUDA
backward kernel (csrc/local_step_bwd.cu) implements and is checked against autograd in tests/.
"""
import math

import torch


def theta_consts_gauss(theta):
    """(W, m, cden) of the Gaussian ELBO denominator log N(x | E[mu], E[Sigma]) + E log pi (svae.py:216-243)."""
    from . import dists
    beta_k, m_k, C_k, v_k = dists.niw_natural_to_standard(*theta[1:])
    mu, sigma = dists.niw_expected_values((beta_k, m_k, C_k, v_k))
    e_log_pi = dists.dirichlet_expected_log_pi(dists.dirichlet_natural_to_standard(theta[0]))
    Ls = torch.linalg.cholesky(sigma)
    D = mu.shape[1]
    W = torch.linalg.solve_triangular(Ls, torch.eye(D, dtype=mu.dtype).expand_as(Ls), upper=False)
    cden = e_log_pi - torch.log(torch.diagonal(Ls, dim1=-2, dim2=-1)).sum(-1) - 0.5 * D * math.log(2 * math.pi)
    return W, mu, cden


def theta_consts_student(theta):
    """(W, m, cden, nu) of the Student-t denominator (svae.py:283-306, student_t.py:39)."""
    from . import dists, svae_port
    mu, sigma = svae_port.unpack_smm(theta[1:3])
    nu = theta[3]
    e_log_pi = dists.dirichlet_expected_log_pi(dists.dirichlet_natural_to_standard(theta[0]))
    Ls = torch.linalg.cholesky(sigma)
    D = mu.shape[1]
    W = torch.linalg.solve_triangular(Ls, torch.eye(D, dtype=mu.dtype).expand_as(Ls), upper=False)
    cden = e_log_pi + torch.lgamma((nu + D) / 2.0) - torch.lgamma(nu / 2.0) - 0.5 * D * torch.log(nu * math.pi) \
        - torch.log(torch.diagonal(Ls, dim1=-2, dim2=-1)).sum(-1)
    return W, mu, cden, nu


def _unpack_phi(eta1_phi2, L_raw, pi_raw):
    L2 = torch.tril(L_raw, -1) + torch.diag_embed(torch.nn.functional.softplus(torch.diagonal(L_raw, dim1=-2, dim2=-1)))
    return L2, L2 @ L2.transpose(-1, -2), torch.log_softmax(pi_raw, dim=-1)


def forward(eta1, eta2d, eta1_phi2, L_raw, pi_raw, W, m, cden, noise, nu=None):
    """eta1, eta2d [N,D]; phi_gmm raw (eta1_phi2[K,D], L_raw[K,D,D], pi_raw[K]); theta record parts W[K,D,D] (lower),
    m[K,D], cden[K] (constants, no gradient); noise[N,K,D,S] -> x_k[N,K,S,D], log_r[N,K], reg (scalar).
    nu[K] selects the Student-t denominator of compute_elbo_smm (svae.py:265-322): cden - (nu+D)/2 log1p(|W(x-m)|^2/nu)."""
    N, D = eta1.shape
    K = eta1_phi2.shape[0]
    S = noise.shape[-1]
    L2, P2, logpi = _unpack_phi(eta1_phi2, L_raw, pi_raw)
    p1 = -2.0 * eta2d
    Pt = P2.unsqueeze(0) + torch.diag_embed(p1).unsqueeze(1)
    L = torch.linalg.cholesky(Pt)
    eta_t = (eta1.unsqueeze(1) + eta1_phi2.unsqueeze(0)).unsqueeze(-1)
    mu_t = torch.cholesky_solve(eta_t, L)                                              # N,K,D,1
    x = (mu_t + torch.linalg.solve_triangular(L.transpose(-1, -2), noise, upper=True)).permute(0, 1, 3, 2)
    mu1 = eta1 / p1
    mu2 = torch.cholesky_solve(eta1_phi2.unsqueeze(-1), L2).squeeze(-1)
    d = mu1.unsqueeze(1) - mu2.unsqueeze(0)                                            # N,K,D
    g = torch.einsum('kij,nkj->nki', P2, d)
    b = torch.cholesky_solve(g.unsqueeze(-1), L).squeeze(-1)
    q = (d * g).sum(-1) - (g * b).sum(-1)
    hld = torch.log(torch.diagonal(L, dim1=-2, dim2=-1)).sum(-1)
    hld2 = torch.log(torch.diagonal(L2, dim1=-2, dim2=-1)).sum(-1)
    score = logpi.unsqueeze(0) - 0.5 * q + hld2.unsqueeze(0) - hld
    log_r = torch.log_softmax(score, dim=1)
    r = torch.exp(log_r)
    e2 = (noise * noise).sum(2)                                                        # N,K,S
    num = -0.5 * e2 + hld.unsqueeze(-1) - 0.5 * D * math.log(2 * math.pi) + log_r.unsqueeze(-1)
    Wx = torch.einsum('kij,nksj->nksi', W, x - m.unsqueeze(0).unsqueeze(2))
    den = _den(cden, (Wx * Wx).sum(-1), nu, D)
    reg = (r.unsqueeze(-1) * (num - den)).sum((0, 1)).mean()
    return x, log_r, reg


def _den(cden, q2, nu, D):
    K = cden.shape[0]
    if nu is None:
        return cden.reshape(1, K, 1) - 0.5 * q2
    nu = nu.reshape(1, K, 1)
    return cden.reshape(1, K, 1) - 0.5 * (nu + D) * torch.log1p(q2 / nu)


def backward_closed_form(eta1, eta2d, eta1_phi2, L_raw, pi_raw, W, m, cden, noise, gx, glr, greg, nu=None):
    """Gradients of  sum(gx * x) + sum(glr * log_r) + greg * reg  w.r.t. (eta1, eta2d, eta1_phi2, L_raw, pi_raw),
    written as the per-pair formulas of DESIGN.md §8 (no autograd)."""
    N, D = eta1.shape
    K = eta1_phi2.shape[0]
    S = noise.shape[-1]
    L2, P2, logpi = _unpack_phi(eta1_phi2, L_raw, pi_raw)
    p1 = -2.0 * eta2d
    Pt = P2.unsqueeze(0) + torch.diag_embed(p1).unsqueeze(1)
    L = torch.linalg.cholesky(Pt)
    Sig = torch.cholesky_inverse(L)                                                    # P~^-1, N,K,D,D
    eta_t = eta1.unsqueeze(1) + eta1_phi2.unsqueeze(0)
    mu_t = torch.einsum('nkij,nkj->nki', Sig, eta_t)
    u = torch.linalg.solve_triangular(L.transpose(-1, -2), noise, upper=True)          # N,K,D,S  (L^-T eps)
    x = (mu_t.unsqueeze(-1) + u).permute(0, 1, 3, 2)
    mu1 = eta1 / p1
    P2inv = torch.cholesky_inverse(L2)
    mu2 = torch.einsum('kij,kj->ki', P2inv, eta1_phi2)
    d = mu1.unsqueeze(1) - mu2.unsqueeze(0)
    g = torch.einsum('kij,nkj->nki', P2, d)
    b = torch.einsum('nkij,nkj->nki', Sig, g)
    q = (d * g).sum(-1) - (g * b).sum(-1)
    hld = torch.log(torch.diagonal(L, dim1=-2, dim2=-1)).sum(-1)
    hld2 = torch.log(torch.diagonal(L2, dim1=-2, dim2=-1)).sum(-1)
    score = logpi.unsqueeze(0) - 0.5 * q + hld2.unsqueeze(0) - hld
    log_r = torch.log_softmax(score, dim=1)
    r = torch.exp(log_r)
    e2 = (noise * noise).sum(2)
    xm = x - m.unsqueeze(0).unsqueeze(2)
    Wx = torch.einsum('kij,nksj->nksi', W, xm)
    q2 = (Wx * Wx).sum(-1)
    den = _den(cden, q2, nu, D)
    coef = torch.ones_like(q2) if nu is None else (nu.reshape(1, K, 1) + D) / (nu.reshape(1, K, 1) + q2)
    num = -0.5 * e2 + hld.unsqueeze(-1) - 0.5 * D * math.log(2 * math.pi) + log_r.unsqueeze(-1)
    T = (num - den).mean(-1)                                                           # N,K
    # ---- upstream of each pair
    Gx = gx + (greg / S) * (r.unsqueeze(-1) * coef).unsqueeze(-1) * torch.einsum('kji,nksj->nksi', W, Wx)
    glr_tot = glr + greg * r * (T + 1.0)
    s_bar = glr_tot - r * glr_tot.sum(1, keepdim=True)                                 # through log_softmax
    hld_bar = greg * r - s_bar                                                         # hld enters reg (+r) and the score (-1)
    # ---- mean path  mu~ = Sig eta~
    gmu = Gx.sum(2)                                                                    # N,K,D
    v = torch.einsum('nkij,nkj->nki', Sig, gmu)                                        # d/d eta~
    Pt_bar = -0.5 * (v.unsqueeze(-1) * mu_t.unsqueeze(-2) + mu_t.unsqueeze(-1) * v.unsqueeze(-2))
    # ---- noise path  u_s = L^-T eps_s :  L_bar = -tril( sum_s u_s t_s^T ),  t_s = L^-1 Gx_s ; then Cholesky reverse
    t = torch.linalg.solve_triangular(L, Gx.permute(0, 1, 3, 2), upper=False)          # N,K,D,S
    L_bar = -torch.tril(torch.einsum('nkis,nkjs->nkij', u, t))
    Phi = torch.tril(L.transpose(-1, -2) @ L_bar)
    Phi = Phi - 0.5 * torch.diag_embed(torch.diagonal(Phi, dim1=-2, dim2=-1))
    Sm = torch.linalg.solve_triangular(L.transpose(-1, -2), torch.linalg.solve_triangular(
        L.transpose(-1, -2), Phi.transpose(-1, -2), upper=True).transpose(-1, -2), upper=True)      # L^-T Phi L^-1
    Pt_bar = Pt_bar + 0.5 * (Sm + Sm.transpose(-1, -2))
    # ---- log-det and the quadratic form of the score
    Pt_bar = Pt_bar + (0.5 * hld_bar).unsqueeze(-1).unsqueeze(-1) * Sig
    Pt_bar = Pt_bar + (-0.5 * s_bar).unsqueeze(-1).unsqueeze(-1) * (b.unsqueeze(-1) * b.unsqueeze(-2))
    dmb = d - b
    d_bar = -s_bar.unsqueeze(-1) * torch.einsum('kij,nkj->nki', P2, dmb)               # ds/dd = -P2 (d - b)
    P2_bar = (-0.5 * s_bar).unsqueeze(-1).unsqueeze(-1) * (d.unsqueeze(-1) * d.unsqueeze(-2)
                                                           - d.unsqueeze(-1) * b.unsqueeze(-2) - b.unsqueeze(-1) * d.unsqueeze(-2))
    P2_bar = (P2_bar + Pt_bar).sum(0)                                                  # K,D,D  (P~ = P2 + diag p1)
    s_sum = s_bar.sum(0)                                                               # K
    P2_bar = P2_bar + (0.5 * s_sum).unsqueeze(-1).unsqueeze(-1) * P2inv                # 1/2 logdet P2 in the score
    p1_bar = torch.diagonal(Pt_bar, dim1=-2, dim2=-1).sum(1)                           # N,D
    # ---- d = mu1 - mu2
    mu1_bar = d_bar.sum(1)
    mu2_bar = -d_bar.sum(0)
    h2_bar = v.sum(0) + torch.einsum('kij,kj->ki', P2inv, mu2_bar)
    w2 = torch.einsum('kij,kj->ki', P2inv, mu2_bar)
    P2_bar = P2_bar - 0.5 * (w2.unsqueeze(-1) * mu2.unsqueeze(-2) + mu2.unsqueeze(-1) * w2.unsqueeze(-2))
    eta1_bar = v.sum(1) + mu1_bar / p1
    p1_bar = p1_bar - mu1_bar * eta1 / (p1 * p1)
    eta2d_bar = -2.0 * p1_bar
    # ---- phi_gmm raw parameters
    L2_bar = torch.tril((P2_bar + P2_bar.transpose(-1, -2)) @ L2)
    dg = torch.diagonal(L_raw, dim1=-2, dim2=-1)
    L_raw_bar = torch.tril(L2_bar, -1) + torch.diag_embed(torch.diagonal(L2_bar, dim1=-2, dim2=-1) * torch.sigmoid(dg))
    pi_raw_bar = s_sum - torch.softmax(pi_raw, dim=-1) * s_sum.sum()
    return eta1_bar, eta2d_bar, h2_bar, L_raw_bar, pi_raw_bar
