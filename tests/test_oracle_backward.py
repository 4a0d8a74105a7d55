"""The hand-derived reverse pass of the local step (oracle/backward.py::backward_closed_form — the specification of
csrc/local_step_bwd.cu) against torch.autograd through the differentiable forward, CPU fp64."""
import pytest
import torch

from oracle import backward as ob


def make_case(N, K, D, S, seed, student=False, dtype=torch.float64):
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=dtype)
    eta2d = -0.5 * (0.5 + torch.rand(N, D, generator=g, dtype=dtype) * 2.0)
    eta1 = rn(N, D)
    eta1_phi2 = rn(K, D)
    L_raw = 0.4 * rn(K, D, D) + torch.eye(D, dtype=dtype)
    pi_raw = rn(K)
    W = torch.tril(0.5 * rn(K, D, D)) + torch.eye(D, dtype=dtype)
    m = rn(K, D)
    cden = rn(K)
    nu = (3.0 + 4.0 * torch.rand(K, generator=g, dtype=dtype)) if student else None
    noise = rn(N, K, D, S)
    gx = rn(N, K, S, D)
    glr = rn(N, K)
    greg = float(rn(1))
    return dict(eta1=eta1, eta2d=eta2d, eta1_phi2=eta1_phi2, L_raw=L_raw, pi_raw=pi_raw, W=W, m=m, cden=cden, nu=nu,
                noise=noise, gx=gx, glr=glr, greg=greg)


def autograd_grads(c):
    leaves = [c[k].clone().requires_grad_(True) for k in ('eta1', 'eta2d', 'eta1_phi2', 'L_raw', 'pi_raw')]
    x, log_r, reg = ob.forward(*leaves, c['W'], c['m'], c['cden'], c['noise'], nu=c['nu'])
    obj = (c['gx'] * x).sum() + (c['glr'] * log_r).sum() + c['greg'] * reg
    return torch.autograd.grad(obj, leaves)


@pytest.mark.parametrize('shape', [(7, 3, 2, 4), (5, 4, 6, 2), (3, 2, 8, 1), (4, 1, 3, 2), (2, 5, 1, 3)])
@pytest.mark.parametrize('student', [False, True])
def test_closed_form_matches_autograd(shape, student):
    c = make_case(*shape, seed=sum(shape), student=student)
    ref = autograd_grads(c)
    got = ob.backward_closed_form(c['eta1'], c['eta2d'], c['eta1_phi2'], c['L_raw'], c['pi_raw'], c['W'], c['m'],
                                  c['cden'], c['noise'], c['gx'], c['glr'], c['greg'], nu=c['nu'])
    for name, a, b in zip(('eta1', 'eta2d', 'eta1_phi2', 'L_raw', 'pi_raw'), got, ref):
        err = (a - b).abs().max() / b.abs().max().clamp_min(1e-30)
        assert err < 1e-10, (name, float(err))


def test_forward_matches_pinned_oracle():
    """The differentiable forward is the same function as the pinned oracle's e_step + compute_elbo regulariser
    (oracle/svae_port.py, itself pinned to the reference's goldens)."""
    import numpy as np
    from oracle import svae_port as sp
    rs = np.random.RandomState(5)
    N, K, D, S = 23, 4, 3, 5
    T = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float64)
    prior, theta = sp.init_mm(K, D, uniform=T(rs.rand(K, D)))
    star0 = sp.m_step(prior, T(2.0 * rs.randn(3 * K + 5, D)), T(rs.dirichlet(np.ones(K), 3 * K + 5)))
    sp.update_gmm_params(theta, star0, 0.5)
    phi_gmm = sp.init_recognition_params(theta, K, normal=T(rs.randn(K)))
    phi_gmm = tuple(t + 0.1 * T(rs.randn(*t.shape)) for t in phi_gmm)
    p1 = np.logaddexp(0.0, rs.randn(N, D))
    eta1, eta2d = T(rs.randn(N, D) * p1), T(-0.5 * p1)
    noise = T(rs.randn(N, K, D, S))
    x_k, log_z, phi_tilde, _ = sp.e_step((eta1, eta2d), phi_gmm, noise)
    y = torch.zeros(N, 2, dtype=torch.float64)
    rec = (torch.zeros(N, K, S, 2, dtype=torch.float64), torch.ones(N, K, S, 2, dtype=torch.float64))
    _, (_, _, _, reg) = sp.compute_elbo(y, rec, theta, phi_tilde, x_k, log_z, 'standard')
    W, m, cden = ob.theta_consts_gauss(theta)
    x2, lr2, reg2 = ob.forward(eta1, eta2d, phi_gmm[0], phi_gmm[1], phi_gmm[2], W, m, cden, noise)
    assert torch.allclose(x2, x_k, rtol=1e-9, atol=1e-10)
    assert torch.allclose(lr2, log_z, rtol=1e-9, atol=1e-10)
    assert abs(float(reg2 - reg)) <= 1e-9 * abs(float(reg))
    # Student-t denominator (compute_elbo_smm)
    th_s = (theta[0], T(rs.randn(K, D)), T(0.3 * rs.randn(K, D, D)) + torch.eye(D, dtype=torch.float64), T(3.0 + rs.rand(K) * 5))
    _, (_, _, _, reg_s) = sp.compute_elbo_smm(y, rec, th_s, phi_tilde, x_k, log_z, 'standard')
    W, m, cden, nu = ob.theta_consts_student(th_s)
    _, _, reg_s2 = ob.forward(eta1, eta2d, phi_gmm[0], phi_gmm[1], phi_gmm[2], W, m, cden, noise, nu=nu)
    assert abs(float(reg_s2 - reg_s)) <= 1e-9 * abs(float(reg_s))
