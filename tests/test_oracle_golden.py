"""Pin the oracle (oracle/*.py, torch fp64) against golden vectors produced by the reference's
own source run over the TF-op shim (tests/golden/make_golden.py), and against independent
known answers (scipy)."""
import math

import numpy as np
import pytest
import scipy.special
import scipy.stats
import torch

from conftest import SVAE_CASES, T, load_golden, regen_decoder, regen_noise
from oracle import dists, mixtures, svae_port

TOL = dict(rtol=1e-9, atol=1e-9)


def close(a, b, rtol=1e-9, atol=1e-9):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    np.testing.assert_allclose(a, np.asarray(b), rtol=rtol, atol=atol)


def golden_state(g):
    names = ['alpha', 'A', 'b', 'beta', 'v_hat']
    prior = [T(g['prior_' + n]) for n in names]
    theta = [T(g['theta_' + n]) for n in names]
    star = [g['star_' + n] for n in names]
    new = [g['new_' + n] for n in names]
    phi_gmm = (T(g['phi_mu_k']), T(g['phi_L_k']), T(g['phi_pi_k']))
    return prior, theta, star, new, phi_gmm


@pytest.mark.parametrize('case', SVAE_CASES)
def test_svae_step_matches_reference(case):
    g = load_golden(case)
    prior, theta, star, new, phi_gmm = golden_state(g)
    noise, u = regen_noise(g)
    y, rec, decoder = regen_decoder(g)
    phi_enc = (T(g['eta1']), T(g['eta2d']))
    x_k, log_r, phi_tilde, dbg = svae_port.e_step(phi_enc, phi_gmm, T(noise))
    close(log_r, g['log_r'], rtol=1e-8, atol=1e-8)
    if 'x_k' in g:
        close(x_k, g['x_k'], rtol=1e-8, atol=1e-8)
        close(phi_tilde[0], g['eta1_tilde']); close(phi_tilde[1], g['eta2_tilde'])
        close(dbg[0], g['w_eta1'], rtol=1e-7, atol=1e-8); close(dbg[1], g['w_eta2'], rtol=1e-7, atol=1e-8)
    else:
        close(x_k[::4], g['x_k_every4'], rtol=1e-8, atol=1e-8)
    xs, z = svae_port.subsample_x(x_k, log_r, T(u))
    close(xs[:, 0, :], g['x_samples'], rtol=1e-8, atol=1e-8)
    elbo, details = svae_port.compute_elbo(T(y), (T(rec[0]), T(rec[1])), theta, phi_tilde, x_k, log_r, decoder)
    close(elbo, g['elbo'], rtol=1e-9, atol=1e-7)
    close(torch.stack(list(details)), g['details'], rtol=1e-9, atol=1e-7)
    st = svae_port.m_step(prior, xs[:, 0, :], torch.exp(log_r))
    for a, b in zip(st, star):
        close(a, b, rtol=1e-9, atol=1e-8)
    svae_port.update_gmm_params(theta, st, float(g['rho']))
    for a, b in zip(theta, new):
        close(a, b, rtol=1e-9, atol=1e-8)


def test_svae_init_matches_reference():
    g = load_golden('svae_init')
    K, D = int(g['K']), int(g['D'])
    prior, theta = svae_port.init_mm(K, D, uniform=T(g['uniform_theta']))
    for i, n in enumerate(['alpha', 'A', 'b', 'beta', 'v_hat']):
        close(prior[i], g['prior_' + n]); close(theta[i], g['theta_' + n])
    mu_k, L_k, pi_k = svae_port.init_recognition_params(theta, K, normal=T(g['normal_pi']))
    close(mu_k, g['phi_mu_k']); close(L_k, g['phi_L_k']); close(pi_k, g['phi_pi_k'])


def test_svae_smm_matches_reference():
    g = load_golden('svae_smm')
    N, K, D, S, seed = (int(g[k]) for k in ('N', 'K', 'D', 'S', 'seed'))
    noise = np.random.RandomState(seed).standard_normal((N, K, D, S))
    y, rec, _ = regen_decoder(g)
    phi_gmm = (T(g['phi_mu_k']), T(g['phi_L_k']), T(g['phi_pi_k']))
    x_k, log_r, phi_tilde, _ = svae_port.e_step((T(g['eta1']), T(g['eta2d'])), phi_gmm, T(noise))
    close(log_r, g['log_r'], rtol=1e-8, atol=1e-8); close(x_k, g['x_k'], rtol=1e-8, atol=1e-8)
    theta = (T(g['theta_alpha']), T(g['theta_mu']), T(g['theta_L']), T(g['theta_dof']))
    close(svae_port.unpack_smm(theta[1:3])[1], g['unpacked_sigma'])
    elbo, details = svae_port.compute_elbo_smm(T(y), (T(rec[0]), T(rec[1])), theta, phi_tilde, x_k, log_r, 'standard')
    close(elbo, g['elbo'], atol=1e-7); close(torch.stack(list(details)), g['details'], atol=1e-7)
    a_star = svae_port.m_step_smm([T(g['prior_alpha'])], torch.exp(log_r))
    close(a_star, g['alpha_star'])
    cur = [theta[0].clone()]
    svae_port.update_gmm_params(cur, [a_star], float(g['rho']))
    close(cur[0], g['alpha_new'])


def _prior(K, D):
    return svae_port.init_mm_params(K, D, alpha_scale=0.05 / K, beta_scale=0.5, m_scale=0, C_scale=D + 0.5,
                                    v_init=D + 0.5)


@pytest.mark.parametrize('case', ['gmm_sweep_a', 'gmm_sweep_b'])
def test_gmm_sweep_matches_reference(case):
    g = load_golden(case)
    x = T(g['x']); K = int(g['K']); D = x.shape[1]
    r_new, theta, (x_k, S_k, pi) = mixtures.gmm_sweep(x, T(g['r0']), _prior(K, D))
    close(r_new, g['r_new'], rtol=1e-8, atol=1e-12)
    for a, n in zip(theta, ['alpha_k', 'beta_k', 'm_k', 'C_k', 'v_k']):
        close(a, g[n])
    close(x_k, g['x_k']); close(S_k, g['S_k']); close(pi, g['pi'])
    P = torch.linalg.inv(theta[3])
    r_m, pi_m = mixtures.gmm_e_step_missing_data(x, theta[0], theta[1], theta[2], P, theta[4],
                                                 torch.as_tensor(g['mask']))
    close(r_m, g['r_miss'], rtol=1e-8, atol=1e-12); close(pi_m, g['pi_miss'])


@pytest.mark.parametrize('case', ['smm_sweep_a', 'smm_sweep_b'])
def test_smm_sweep_matches_reference(case):
    g = load_golden(case)
    x = T(g['x']); K = int(g['K']); D = x.shape[1]
    kappa = torch.full((K,), float(g['kappa']), dtype=torch.float64)
    prior = _prior(K, D)
    r1, u1, theta, (x_k, S_k, pi) = mixtures.smm_sweep(x, T(g['r0']), torch.ones_like(T(g['r0'])), prior, kappa)
    close(r1, g['r1'], rtol=1e-8, atol=1e-12); close(u1, g['u1'], rtol=1e-8)
    for a, n in zip(theta[:5], ['alpha_k', 'beta_k', 'm_k', 'C_k', 'v_k']):
        close(a, g[n])
    close(x_k, g['x_k']); close(S_k, g['S_k']); close(pi, g['pi'])
    r2, u2, theta2, (xk2, Sk2, pi2) = mixtures.smm_sweep(x, r1, u1, prior, kappa)
    close(r2, g['r2'], rtol=1e-8, atol=1e-12); close(u2, g['u2'], rtol=1e-8)
    for a, n in zip(theta2[:5], ['alpha2', 'beta2', 'm2', 'C2', 'v2']):
        close(a, g[n])
    close(xk2, g['xk2']); close(Sk2, g['Sk2']); close(pi2, g['pi2'])


def test_distributions_match_reference():
    g = load_golden('distributions')
    e1, e2 = dists.gaussian_standard_to_natural(T(g['mu']), T(g['sigma']))
    close(e1, g['eta1']); close(e2, g['eta2'])
    mu, sg = dists.gaussian_natural_to_standard(e1, e2)
    close(mu, g['mu_back']); close(sg, g['sigma_back'])
    close(dists.gaussian_log_probability_nat(T(g['x']), T(g['eta1_nk']), T(g['eta2_nk']), T(g['w'])), g['logprob_nat'])
    close(dists.gaussian_log_probability_nat(T(g['x']), T(g['eta1_nk']), T(g['eta2_nk'])), g['logprob_nat_noweights'])
    close(dists.gaussian_log_probability_nat_per_samp(T(g['xs']), T(g['eta1_nk']), T(g['eta2_nk'])), g['logprob_per_samp'])
    A, b, beta, vh = dists.niw_standard_to_natural(T(g['beta']), T(g['m']), T(g['C']), T(g['v']))
    close(A, g['A']); close(b, g['b']); close(vh, g['v_hat'])
    back = dists.niw_natural_to_standard(A, b, beta, vh)
    close(back[1], g['back_m']); close(back[2], g['back_C']); close(back[3], g['back_v'])
    em, eC = dists.niw_expected_values(back)
    close(em, g['exp_m']); close(eC, g['exp_C'])
    close(dists.dirichlet_expected_log_pi(T(g['alpha'])), g['expected_log_pi'])
    close(dists.student_t_log_probability_per_samp(T(g['xs']), T(g['mu']), T(g['sigma']), T(g['dof'])), g['student_per_samp'])
    close(dists.student_t_logprob_smm_mixture(T(g['x']), T(g['mu']), T(g['sigma']), T(g['dof']), torch.log(T(g['w']))),
          g['student_mixture'])


# ------------------------------------------------------------------ independent known answers
def test_kat_gaussian_logprob_vs_scipy():
    rs = np.random.RandomState(0)
    N, K, D = 5, 3, 4
    mu = rs.randn(K, D); a = rs.randn(K, D, D); sigma = a @ a.transpose(0, 2, 1) + D * np.eye(D)
    x = rs.randn(N, D); w = rs.dirichlet(np.ones(K))
    e1, e2 = dists.gaussian_standard_to_natural(T(mu), T(sigma))
    lp = dists.gaussian_log_probability_nat(T(x), e1.unsqueeze(0).expand(N, K, D), e2.unsqueeze(0).expand(N, K, D, D), T(w))
    ref = np.stack([scipy.stats.multivariate_normal.logpdf(x, mu[k], sigma[k]) + np.log(w[k]) for k in range(K)], 1)
    ref = ref - scipy.special.logsumexp(ref, axis=1, keepdims=True)
    close(lp, ref, rtol=1e-10, atol=1e-10)
    xs = rs.randn(N, K, 2, D)
    lps = dists.gaussian_log_probability_nat_per_samp(T(xs), e1.unsqueeze(0).expand(N, K, D).contiguous(),
                                                      e2.unsqueeze(0).expand(N, K, D, D).contiguous())
    ref = np.stack([scipy.stats.multivariate_normal.logpdf(xs[:, k], mu[k], sigma[k]) for k in range(K)], 1)
    close(lps, ref, rtol=1e-10, atol=1e-10)


def test_kat_student_t_vs_scipy():
    rs = np.random.RandomState(1)
    N, K, S, D = 4, 3, 2, 3
    mu = rs.randn(K, D); a = rs.randn(K, D, D); sigma = a @ a.transpose(0, 2, 1) + np.eye(D)
    dof = np.array([2.5, 5.0, 30.0]); y = rs.randn(N, K, S, D)
    lp = dists.student_t_log_probability_per_samp(T(y), T(mu), T(sigma), T(dof))
    ref = np.stack([scipy.stats.multivariate_t.logpdf(y[:, k], mu[k], sigma[k], df=dof[k]) for k in range(K)], 1)
    close(lp, ref, rtol=1e-10, atol=1e-10)


def test_kat_dirichlet_and_roundtrips():
    alpha = np.array([0.3, 1.0, 2.5, 7.0])
    close(dists.dirichlet_expected_log_pi(T(alpha)), scipy.special.digamma(alpha) - scipy.special.digamma(alpha.sum()))
    close(dists.dirichlet_natural_to_standard(dists.dirichlet_standard_to_natural(T(alpha))), alpha)
    rs = np.random.RandomState(2)
    K, D = 3, 4
    beta, m, v = rs.rand(K) + 0.2, rs.randn(K, D), rs.rand(K) + D + 2
    a = rs.randn(K, D, D); C = a @ a.transpose(0, 2, 1) + np.eye(D)
    back = dists.niw_natural_to_standard(*dists.niw_standard_to_natural(T(beta), T(m), T(C), T(v)))
    for x, y in zip(back, (beta, m, C, v)):
        close(x, y, rtol=1e-12, atol=1e-12)
    em, eC = dists.niw_expected_values(back)
    close(eC, C / v[:, None, None], rtol=1e-10)          # E[Sigma] = C / v


def test_kat_single_component_and_1d():
    # K = 1  =>  log r = 0 ; D = 1 closed form for the e-step score
    rs = np.random.RandomState(3)
    N, D = 6, 3
    eta2d = -0.5 * np.logaddexp(0, rs.randn(N, D)); eta1 = rs.randn(N, D)
    phi = (T(rs.randn(1, D)), T(rs.randn(1, D, D)), T(rs.randn(1)))
    _, log_r, _, _ = svae_port.e_step((T(eta1), T(eta2d)), phi, T(rs.randn(N, 1, D, 2)))
    close(log_r, np.zeros((N, 1)), atol=1e-12)
    K = 3
    eta2d = -0.5 * np.logaddexp(0, rs.randn(N, 1)); eta1 = rs.randn(N, 1)
    h2, Lr, pr = rs.randn(K, 1), rs.randn(K, 1, 1), rs.randn(K)
    _, log_r, _, _ = svae_port.e_step((T(eta1), T(eta2d)), (T(h2), T(Lr), T(pr)), T(rs.randn(N, K, 1, 1)))
    p1 = -2 * eta2d[:, 0]; mu1 = eta1[:, 0] / p1
    p2 = np.logaddexp(0, Lr[:, 0, 0]) ** 2; mu2 = h2[:, 0] / p2
    var = 1 / p1[:, None] + 1 / p2[None, :]
    s = scipy.stats.norm.logpdf(mu1[:, None], mu2[None, :], np.sqrt(var)) + np.log(scipy.special.softmax(pr))[None]
    close(log_r, s - scipy.special.logsumexp(s, axis=1, keepdims=True), rtol=1e-10, atol=1e-10)


def test_kat_bishop_m_step_tiny():
    # Bishop 10.51-10.63 by hand for N=2, K=1, D=1, including the reference's +1 on v_k (gmm.py:81)
    x = T([[1.0], [3.0]]); r = T([[0.5], [1.0]])
    a0, b0, m0, C0, v0 = T([0.1]), T([2.0]), T([[0.5]]), T([[[1.5]]]), T([4.0])
    alpha, beta, m, C, v, xk, Sk = mixtures.gmm_m_step(x, r, a0, b0, m0, C0, v0)
    Nk = 1.5; xbar = (0.5 * 1 + 3) / 1.5; S = (0.5 * (1 - xbar) ** 2 + (3 - xbar) ** 2) / 1.5
    close(alpha, [0.1 + Nk]); close(beta, [2 + Nk]); close(m, [[(2 * 0.5 + Nk * xbar) / (2 + Nk)]])
    close(C, [[[1.5 + Nk * S + 2 * Nk / (2 + Nk) * (xbar - 0.5) ** 2]]]); close(v, [4 + Nk + 1])


def test_m_step_is_additive_in_natural_parameters():
    """SURVEY 8a-note 4 (what the kernels rely on): theta* = prior + [N_k, sum r x x^T, sum r x, N_k, N_k + 1]."""
    rs = np.random.RandomState(4)
    N, K, D = 50, 4, 3
    prior, _ = svae_port.init_mm(K, D, uniform=T(rs.rand(K, D)))
    x, r = T(rs.randn(N, D) * 2), T(rs.dirichlet(np.ones(K), N))
    r[:, 2] = 0.0                                   # an empty component exercises the NaN guards
    star = svae_port.m_step(prior, x, r)
    Nk = r.sum(0)
    close(star[0], prior[0] + Nk); close(star[3], prior[3] + Nk); close(star[4], prior[4] + Nk + 1)
    close(star[2], prior[2] + torch.einsum('nk,nd->kd', r, x), rtol=1e-12, atol=1e-12)
    close(star[1], prior[1] + torch.einsum('nk,nd,ne->kde', r, x, x), rtol=1e-11, atol=1e-11)


def test_multinomial_inverse_cdf_semantics():
    logits = torch.log(T([[0.1, 0.2, 0.3, 0.4]]))
    for u, z in ((0.0, 0), (0.0999, 0), (0.1001, 1), (0.5999, 2), (0.6001, 3), (0.99999, 3)):
        assert int(svae_port.multinomial_inverse_cdf(logits, T([[u]]))[0, 0]) == z


def test_cvi_step_size():
    assert math.isclose(svae_port.cvi_step_size(0.2, 2500, 0.95), 0.2 * 0.95 ** 2.5)
