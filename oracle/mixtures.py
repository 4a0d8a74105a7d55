"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's
Bayesian GMM (models/gmm.py, Bishop PRML 10.2) and Student-t mixture (models/smm.py,
Archambeau & Verleysen 2007) VB-EM building blocks.  torch-CPU, dtype-generic.
Pinned against the reference's own source run over oracle/tf_shim (tests/golden).
Quirks are kept verbatim: `+1` in gmm update_vk, `psi((v+1+i)/2)` vs `psi((v+i)/2)`,
the det>1e-20 guard, the linear-in-distance SMM responsibilities.
"""
import math

import torch

from . import dists


def _isnan_where(normed, raw):
    return torch.where(torch.isnan(normed), raw, normed)


# ======================================================================= GMM (models/gmm.py)
def gmm_update_Nk(r_nk):
    """gmm.py:25-27."""
    return r_nk.sum(0)


def gmm_update_xk(x, r_nk, N_k):
    """gmm.py:30-36 (NaN -> unnormalised when N_k == 0)."""
    x_k = torch.einsum('nk,nd->kd', r_nk, x)
    return _isnan_where(x_k / N_k.unsqueeze(1), x_k)


def gmm_update_Sk(x, r_nk, N_k, x_k):
    """gmm.py:39-46."""
    x_xk = x.unsqueeze(1) - x_k.unsqueeze(0)
    S = torch.einsum('nk,nkde->kde', r_nk, torch.einsum('nkd,nke->nkde', x_xk, x_xk))
    return _isnan_where(S / N_k.unsqueeze(1).unsqueeze(2), S)


def gmm_update_mk(beta_0, m_0, N_k, x_k, beta_k):
    """gmm.py:59-67."""
    if beta_0.dim() == 1:
        beta_0 = beta_0.reshape(-1, 1)
    return (beta_0 * m_0 + N_k.unsqueeze(1) * x_k) / beta_k.unsqueeze(1)


def gmm_update_Ck(C_0, x_k, N_k, m_0, beta_0, beta_k, S_k):
    """gmm.py:70-76."""
    C = C_0 + N_k.unsqueeze(1).unsqueeze(2) * S_k
    Q0 = x_k - m_0
    q = torch.einsum('kd,ke->kde', Q0, Q0)
    return C + torch.einsum('k,kde->kde', beta_0 * N_k / beta_k, q)


def gmm_m_step(x, r_nk, alpha_0, beta_0, m_0, C_0, v_0):
    """gmm.py:201-227."""
    N_k = gmm_update_Nk(r_nk)
    x_k = gmm_update_xk(x, r_nk, N_k)
    S_k = gmm_update_Sk(x, r_nk, N_k, x_k)
    alpha_k = alpha_0 + N_k                                   # gmm.py:49-51
    beta_k = beta_0 + N_k                                     # gmm.py:54-56
    m_k = gmm_update_mk(beta_0, m_0, N_k, x_k, beta_k)
    C_k = gmm_update_Ck(C_0, x_k, N_k, m_0, beta_0, beta_k, S_k)
    v_k = v_0 + N_k + 1                                       # gmm.py:79-81 (the +1 quirk)
    return alpha_k, beta_k, m_k, C_k, v_k, x_k, S_k


def expct_mahalanobis_dist(x, beta_k, m_k, P_k, v_k):
    """gmm.py:84-94 == smm.py:88-97."""
    D = x.shape[1]
    dist = x.unsqueeze(1) - m_k.unsqueeze(0)
    m = torch.einsum('k,nk->nk', v_k,
                     torch.einsum('nkd,nkd->nk', dist, torch.einsum('kde,nke->nkd', P_k, dist)))
    return m + (D / beta_k).reshape(1, -1)


def gmm_compute_dev_missing_data(x, beta_k, m_k, P_k, v_k, missing_data_mask):
    """gmm.py:97-114."""
    D = x.shape[1]
    d_beta = (D / beta_k).reshape(1, -1)
    x_mk = x.unsqueeze(1) - m_k.unsqueeze(0)
    x_mk = x_mk * (~missing_data_mask).to(x.dtype).unsqueeze(1)
    m = torch.einsum('k,nk->nk', v_k,
                     torch.einsum('nkd,nkd->nk', x_mk, torch.einsum('kde,nke->nkd', P_k, x_mk)))
    return d_beta + m


def gmm_compute_expct_log_det_prec(v_k, P_k):
    """gmm.py:117-131 (det-threshold guard; psi((v+1+i)/2), i=0..D-1)."""
    det_P = torch.linalg.det(P_k)
    log_det_P = torch.where(det_P > 1e-20, torch.log(det_P), torch.zeros_like(det_P))
    D = P_k.shape[1]
    i = torch.arange(D, dtype=P_k.dtype).unsqueeze(0)
    sum_digamma = torch.special.digamma(0.5 * (v_k.unsqueeze(1) + 1.0 + i)).sum(1)
    return sum_digamma + D * math.log(2.0) + log_det_P


def compute_log_pi(alpha_k):
    """gmm.py:134-138 == smm.py:113-116."""
    return torch.special.digamma(alpha_k) - torch.special.digamma(alpha_k.sum())


def gmm_compute_rnk(expct_log_pi, expct_log_det_cov, expct_dev):
    """gmm.py:141-151."""
    log_rho = expct_log_pi + 0.5 * expct_log_det_cov - 0.5 * expct_dev
    rho = torch.exp(log_rho - log_rho.max(dim=1).values.reshape(-1, 1))
    return rho / rho.sum(1).unsqueeze(1)


def gmm_e_step(x, alpha_k, beta_k, m_k, P_k, v_k):
    """gmm.py:154-174 -> (r_nk, exp(E log pi))."""
    dev = expct_mahalanobis_dist(x, beta_k, m_k, P_k, v_k)
    ldc = gmm_compute_expct_log_det_prec(v_k, P_k)
    lpi = compute_log_pi(alpha_k)
    return gmm_compute_rnk(lpi, ldc, dev), torch.exp(lpi)


def gmm_e_step_missing_data(x, alpha_k, beta_k, m_k, P_k, v_k, missing_data_mask):
    """gmm.py:177-198."""
    dev = gmm_compute_dev_missing_data(x, beta_k, m_k, P_k, v_k, missing_data_mask)
    ldc = gmm_compute_expct_log_det_prec(v_k, P_k)
    lpi = compute_log_pi(alpha_k)
    return gmm_compute_rnk(lpi, ldc, dev), torch.exp(lpi)


def gmm_sweep(x, r_nk, prior_nat):
    """gmm.py:230-269 (`inference`) for a given state r_nk and Dirichlet+NIW prior in
    natural parameters: m_step -> P = inv(C) -> e_step.  Returns (r_new, theta, (x_k,S_k,pi))."""
    alpha, A, b, beta, v_hat = prior_nat
    beta_0, m_0, C_0, v_0 = dists.niw_natural_to_standard(A, b, beta, v_hat)
    alpha_0 = dists.dirichlet_natural_to_standard(alpha)
    alpha_k, beta_k, m_k, C_k, v_k, x_k, S_k = gmm_m_step(x, r_nk, alpha_0, beta_0, m_0, C_0, v_0)
    P_k = torch.linalg.inv(C_k)
    r_new, pi = gmm_e_step(x, alpha_k, beta_k, m_k, P_k, v_k)
    return r_new, (alpha_k, beta_k, m_k, C_k, v_k), (x_k, S_k, pi)


# ======================================================================= SMM (models/smm.py)
def smm_update_xk(x, ru_nk, W_k, eps=1e-20):
    """smm.py:35-40."""
    return torch.einsum('nk,nd->kd', ru_nk, x) / (W_k.unsqueeze(1) + eps)


def smm_update_Sk(x, ru_nk, W_k, x_k, eps=1e-20):
    """smm.py:43-50."""
    err = x.unsqueeze(1) - x_k.unsqueeze(0)
    S_k = torch.einsum('nk,nkde->kde', ru_nk, torch.einsum('nkd,nke->nkde', err, err))
    return S_k / (W_k.unsqueeze(1).unsqueeze(2) + eps)


def smm_m_step(x, r_nk, u_nk, alpha_0, beta_0, m_0, C_0, v_0):
    """smm.py:167-196 (v_k = v_0 + N_k: no +1 here, smm.py:74-76)."""
    ru = r_nk * u_nk
    N_k = r_nk.sum(0)                                          # smm.py:25-27
    W_k = ru.sum(0)                                            # smm.py:30-32
    x_k = smm_update_xk(x, ru, W_k)
    S_k = smm_update_Sk(x, ru, W_k, x_k)
    alpha_k = alpha_0 + N_k                                    # smm.py:53-55
    beta_k = beta_0 + W_k                                      # smm.py:58-60
    m_k = gmm_update_mk(beta_0, m_0, W_k, x_k, beta_k)         # smm.py:63-71 (same formula)
    C_k = gmm_update_Ck(C_0, x_k, W_k, m_0, beta_0, beta_k, S_k)   # smm.py:79-85 (same formula)
    v_k = v_0 + N_k
    return alpha_k, beta_k, m_k, C_k, v_k, x_k, S_k


def smm_expct_log_det_prec(v_k, P_k):
    """smm.py:100-110 (Cholesky logdet; psi((v+i)/2), i=0..D-1)."""
    D = P_k.shape[1]
    i = torch.arange(D, dtype=P_k.dtype).unsqueeze(0)
    return torch.special.digamma(0.5 * (v_k.unsqueeze(1) + i)).sum(1) + D * math.log(2.0) + dists.logdet(P_k)


def smm_compute_rnk(expct_log_pi, expct_log_det_prec, expct_m_dist, kappa_k, D):
    """smm.py:119-128 with the literal operator precedence of line 124."""
    log_r = torch.lgamma((D + kappa_k) / 2.0) - torch.lgamma(kappa_k / 2.0) - (D / 2.0) * torch.log(kappa_k * math.pi)
    log_r = log_r + expct_log_pi + 0.5 * expct_log_det_prec
    log_r = log_r - (0.5 * (D + kappa_k) * expct_m_dist - torch.log(kappa_k))
    return torch.exp(log_r - torch.logsumexp(log_r, dim=1, keepdim=True))


def smm_compute_expct_unk(expct_m_dist, kappa_k, D):
    """smm.py:131-137."""
    return (0.5 * (D + kappa_k)) / (0.5 * (expct_m_dist + kappa_k))


def smm_e_step(x, alpha_k, beta_k, m_k, P_k, v_k, kappa_k):
    """smm.py:140-164 -> (r_nk, u_nk, exp(E log pi))."""
    m_dist = expct_mahalanobis_dist(x, beta_k, m_k, P_k, v_k)
    ldp = smm_expct_log_det_prec(v_k, P_k)
    lpi = compute_log_pi(alpha_k)
    D = x.shape[1]
    return smm_compute_rnk(lpi, ldp, m_dist, kappa_k, D), smm_compute_expct_unk(m_dist, kappa_k, D), torch.exp(lpi)


def smm_sweep(x, r_nk, u_nk, prior_nat, kappa_k):
    """smm.py:199-245 (`inference`) for a given state (r,u): m_step -> P=inv(C) -> e_step."""
    alpha, A, b, beta, v_hat = prior_nat
    beta_0, m_0, C_0, v_0 = dists.niw_natural_to_standard(A, b, beta, v_hat)
    alpha_0 = dists.dirichlet_natural_to_standard(alpha)
    alpha_k, beta_k, m_k, C_k, v_k, x_k, S_k = smm_m_step(x, r_nk, u_nk, alpha_0, beta_0, m_0, C_0, v_0)
    P_k = torch.linalg.inv(C_k)
    r_new, u_new, pi = smm_e_step(x, alpha_k, beta_k, m_k, P_k, v_k, kappa_k)
    return r_new, u_new, (alpha_k, beta_k, m_k, C_k, v_k, kappa_k), (x_k, S_k, pi)
