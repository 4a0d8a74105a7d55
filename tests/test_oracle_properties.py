"""Property tests (hypothesis) of the oracle over random small shapes and positive-definite inputs:
the identities the CUDA kernels rely on against the literal restatement of the reference graph."""
import math

import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import svae_port as sp

T = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float64)


def _inputs(N, K, D, S, seed):
    rs = np.random.RandomState(seed)
    prior, theta = sp.init_mm(K, D, uniform=T(rs.rand(K, D)))
    phi_gmm = sp.init_recognition_params(theta, K, normal=T(rs.randn(K)))
    phi_gmm = (phi_gmm[0] + 0.3 * T(rs.randn(K, D)), phi_gmm[1] + 0.2 * T(rs.randn(K, D, D)), phi_gmm[2] + T(rs.randn(K)))
    p1 = np.logaddexp(0.0, 2.0 * rs.randn(N, D)) + 1e-3
    phi_enc = (T(rs.randn(N, D) * p1), T(-0.5 * p1))
    return prior, theta, phi_gmm, phi_enc, T(rs.randn(N, K, D, S))


shapes = st.tuples(st.integers(1, 6), st.integers(1, 4), st.integers(1, 5), st.integers(1, 3), st.integers(0, 10 ** 6))


@settings(max_examples=25, deadline=None)
@given(shapes)
def test_responsibilities_are_a_distribution_and_fused_identities_hold(sh):
    N, K, D, S, seed = sh
    _, _, phi_gmm, phi_enc, noise = _inputs(N, K, D, S, seed)
    x_k, log_z, (eta1_t, eta2_t), _ = sp.e_step(phi_enc, phi_gmm, noise)
    assert torch.isfinite(log_z).all() and torch.allclose(torch.exp(log_z).sum(1), torch.ones(N, dtype=torch.float64), atol=1e-10)
    if K == 1:
        assert float(log_z.abs().max()) < 1e-10                                   # a single component takes everything
    # note 1 (centred form) and note 2/3 (samples and their log-density through one Cholesky)
    eta1_phi2, eta2_phi2, pi = sp.unpack_recognition_gmm(phi_gmm)
    P2 = -2.0 * eta2_phi2
    p1 = -2.0 * phi_enc[1]
    mu1 = phi_enc[0] / p1
    mu2 = torch.linalg.solve(P2, eta1_phi2.unsqueeze(-1)).squeeze(-1)
    score = torch.empty(N, K, dtype=torch.float64)
    for n in range(N):
        for k in range(K):
            Pt = P2[k] + torch.diag(p1[n])
            L = torch.linalg.cholesky(Pt)
            d = mu1[n] - mu2[k]
            a = torch.linalg.solve_triangular(L, (P2[k] @ d).unsqueeze(-1), upper=False).squeeze(-1)
            a1 = torch.linalg.solve_triangular(L, (p1[n] * d).unsqueeze(-1), upper=False).squeeze(-1)
            score[n, k] = torch.log(pi[k]) - 0.5 * (a @ a1) + 0.5 * torch.logdet(P2[k]) - torch.log(torch.diagonal(L)).sum()
            y = torch.linalg.solve_triangular(L.T, noise[n, k] - a.unsqueeze(-1), upper=True)          # D,S
            assert torch.allclose(mu1[n].unsqueeze(-1) + y, x_k[n, k].T, rtol=1e-8, atol=1e-8)
            lp = -0.5 * (noise[n, k] ** 2).sum(0) + torch.log(torch.diagonal(L)).sum() - 0.5 * D * math.log(2 * math.pi)
            from oracle import dists
            ref = dists.gaussian_log_probability_nat_per_samp(x_k[n:n + 1, k:k + 1], eta1_t[n:n + 1, k:k + 1].reshape(1, 1, D),
                                                              eta2_t[n:n + 1, k:k + 1])
            assert torch.allclose(lp, ref.reshape(-1), rtol=1e-8, atol=1e-8)
    assert torch.allclose(torch.log_softmax(score, dim=1), log_z, rtol=1e-8, atol=1e-8)


@settings(max_examples=25, deadline=None)
@given(st.integers(2, 12), st.integers(1, 4), st.integers(1, 5), st.integers(0, 10 ** 6))
def test_m_step_statistics_are_additive_over_points(N, K, D, seed):
    rs = np.random.RandomState(seed)
    prior, _ = sp.init_mm(K, D, uniform=T(rs.rand(K, D)))
    x, r = T(rs.randn(N, D) * 2), T(rs.dirichlet(np.ones(K), N))
    cut = rs.randint(1, N)
    full = sp.m_step(prior, x, r)
    a, b = sp.m_step(prior, x[:cut], r[:cut]), sp.m_step(prior, x[cut:], r[cut:])
    for f, pa, pb, p0, extra in zip(full, a, b, prior, (0.0, 0.0, 0.0, 0.0, 1.0)):
        # natural parameters: star - prior is a sum over points (+1 on v_hat, the reference's gmm.update_vk quirk)
        assert torch.allclose(f - p0 - extra, (pa - p0 - extra) + (pb - p0 - extra), rtol=1e-9, atol=1e-9)
