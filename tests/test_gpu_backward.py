"""Reverse pass of the fused local step (csrc/local_step_bwd.cu behind vmp_for_svae_b200.autograd) against
torch.autograd through the oracle's differentiable forward (oracle/backward.py, pinned to svae_port in
tests/test_oracle_backward.py) on identical inputs and injected noise."""
import numpy as np
import pytest
import torch

from conftest import T

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'

BWD_SHAPES = [(100, 10, 2, 10), (274, 10, 6, 10), (33, 5, 11, 3), (257, 32, 8, 2), (40, 3, 16, 1), (130, 1, 5, 2),
              (1, 3, 4, 1), (61, 200, 3, 1),
              # D > 16 / K > 256: the block-cooperative kernel (one CTA per point, matrices in shared memory)
              (64, 12, 32, 1), (40, 9, 64, 1), (20, 5, 24, 2), (9, 4, 48, 3), (12, 300, 3, 1)]


def _inputs(N, K, D, S, seed, student):
    from test_gpu_parity import _oracle_inputs
    prior, theta, phi_gmm, phi_enc, noise, _ = _oracle_inputs(N, K, D, S, seed=seed, spread=0.3)
    rs = np.random.RandomState(seed + 1)
    if student:
        theta = (theta[0], T(rs.randn(K, D)), T(0.3 / D ** 0.5 * rs.randn(K, D, D)) + torch.eye(D, dtype=torch.float64),
                 T(3.0 + 5.0 * rs.rand(K)))
    gx = T(rs.randn(N, K, S, D)) * T((rs.rand(N, K, 1, 1) < 0.3).astype(np.float64))     # sparse like the z-gather
    glr, greg = T(rs.randn(N, K)), float(rs.randn())
    return theta, phi_gmm, phi_enc, noise, gx, glr, greg


def _oracle_grads(theta, phi_gmm, phi_enc, noise, gx, glr, greg, student):
    from oracle import backward as ob
    if student:
        W, m, cden, nu = ob.theta_consts_student(theta)
    else:
        (W, m, cden), nu = ob.theta_consts_gauss(theta), None
    leaves = [t.clone().requires_grad_(True) for t in (phi_enc[0], phi_enc[1], phi_gmm[0], phi_gmm[1], phi_gmm[2])]
    x, log_r, reg = ob.forward(*leaves, W, m, cden, noise, nu=nu)
    obj = (gx * x).sum() + (glr * log_r).sum() + greg * reg
    return torch.autograd.grad(obj, leaves), (x.detach(), log_r.detach(), reg.detach())


def _gpu_grads(theta, phi_gmm, phi_enc, noise, gx, glr, greg, student, dt, seed=None):
    from vmp_for_svae_b200 import core
    from vmp_for_svae_b200.autograd import local_step_autograd
    dev = lambda t: t.to(device=DEV, dtype=dt).contiguous()
    th = [dev(t) for t in theta]
    theta_rec = core.theta_prepare_student(th) if student else core.theta_prepare_gauss(th)
    leaves = [dev(t).requires_grad_(True) for t in (phi_enc[0], phi_enc[1], phi_gmm[0], phi_gmm[1], phi_gmm[2])]
    S = gx.shape[2]
    x, log_r, reg, acc = local_step_autograd(*leaves, theta_rec, S, den_mode=core.DEN_STUDENT if student else core.DEN_GAUSS,
                                             noise=None if noise is None else dev(noise), seed=seed or 0)
    obj = (dev(gx) * x).sum() + (dev(glr) * log_r).sum() + greg * reg
    grads = torch.autograd.grad(obj, leaves)
    torch.cuda.synchronize()
    assert float(acc[3]) == 0
    return [g.cpu().double() for g in grads], (x.detach().cpu().double(), log_r.detach().cpu().double(), reg.detach().cpu().double())


@pytest.mark.parametrize('student', [False, True], ids=['gauss', 'student'])
@pytest.mark.parametrize('dt', [torch.float64, torch.float32], ids=['f64', 'f32'])
@pytest.mark.parametrize('shape', BWD_SHAPES, ids=lambda s: 'N%dK%dD%dS%d' % s)
def test_backward_vs_oracle_autograd(shape, dt, student):
    N, K, D, S = shape
    args = _inputs(N, K, D, S, seed=N + 3 * K + D, student=student)
    ref, fref = _oracle_grads(*args, student)
    got, fgot = _gpu_grads(*args, student, dt)
    # tolerance written here: fp64 1e-8, fp32 2e-3 of the largest gradient entry (the Cholesky reverse amplifies rounding
    # by the condition number of P~, as TF's fp32 graph does)
    tol = 1e-8 if dt == torch.float64 else 2e-3
    assert torch.allclose(fgot[1].exp(), fref[1].exp(), rtol=0, atol=1e-9 if dt == torch.float64 else 1e-4)
    for name, a, b in zip(('eta1', 'eta2_diag', 'eta1_phi2', 'L_raw', 'pi_raw'), got, ref):
        assert a.shape == b.shape
        err = float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
        assert err < tol, (name, shape, str(dt), err)
    # structure: only the lower triangle of L_raw receives gradient (svae.py:351 tf.matrix_band_part)
    assert float(torch.triu(got[3], 1).abs().max()) == 0.0


def test_backward_inkernel_noise_equals_injected():
    from vmp_for_svae_b200 import core
    N, K, D, S = 97, 6, 6, 4
    theta, phi_gmm, phi_enc, _, gx, glr, greg = _inputs(N, K, D, S, seed=11, student=False)
    noise, _ = core.fill_noise(N, K, D, S, 4242, torch.float64, DEV, want_u=False)
    a, _ = _gpu_grads(theta, phi_gmm, phi_enc, noise.cpu(), gx, glr, greg, False, torch.float64)
    b, _ = _gpu_grads(theta, phi_gmm, phi_enc, None, gx, glr, greg, False, torch.float64, seed=4242)
    for x, y in zip(a, b):
        assert torch.allclose(x, y, rtol=1e-12, atol=1e-12)


def test_backward_through_selected_sample_and_decoder():
    """End to end as the reference trains: x_samps = x_k[n, z_n] -> a decoder -> loss - reg; gradients of the encoder
    outputs and phi_gmm against the oracle graph (experiments.py:208-232)."""
    from oracle import backward as ob
    from vmp_for_svae_b200 import core
    from vmp_for_svae_b200.autograd import local_step_autograd
    N, K, D, S = 120, 10, 6, 3
    theta, phi_gmm, phi_enc, noise, _, _, _ = _inputs(N, K, D, S, seed=3, student=False)
    rs = np.random.RandomState(9)
    z = torch.as_tensor(rs.randint(0, K, N))
    Wd, y = T(rs.randn(D, 5)), T(rs.randn(N, 5))

    def loss_of(x, log_r, reg):
        xs = x[torch.arange(N, device=x.device), z.to(x.device)]                 # N,S,D
        rec = torch.tanh(xs @ Wd.to(x)) - y.to(x).unsqueeze(1)
        return 0.5 * (rec * rec).sum() / S + reg
    W, m, cden = ob.theta_consts_gauss(theta)
    lv = [t.clone().requires_grad_(True) for t in (phi_enc[0], phi_enc[1], phi_gmm[0], phi_gmm[1], phi_gmm[2])]
    ref = torch.autograd.grad(loss_of(*ob.forward(*lv, W, m, cden, noise)), lv)
    dev = lambda t: t.to(device=DEV).contiguous()
    theta_rec = core.theta_prepare_gauss([dev(t) for t in theta])
    lg = [dev(t).requires_grad_(True) for t in (phi_enc[0], phi_enc[1], phi_gmm[0], phi_gmm[1], phi_gmm[2])]
    x, log_r, reg, _ = local_step_autograd(*lg, theta_rec, S, noise=dev(noise))
    got = torch.autograd.grad(loss_of(x, log_r, reg), lg)
    for a, b in zip(got, ref):
        assert float((a.cpu() - b).abs().max() / b.abs().max()) < 1e-8


def test_backward_argument_errors():
    from vmp_for_svae_b200 import core
    N, K, D, S = 4, 2, 65, 1            # latent dimensions up to 64 are supported (block-cooperative kernel above 16)
    z = lambda *s: torch.zeros(*s, dtype=torch.float64, device=DEV)
    plen, tlen, _ = __import__('vmp_for_svae_b200')._lib.record_lens(D)
    with pytest.raises(ValueError):
        core.local_step_backward(z(N, D), z(N, D), z(K, D), z(K, D, D), z(K), z(K, plen), z(K, tlen), S, z(N, K),
                                 z(N, K, S, D), z(N, K), 1.0)
    with pytest.raises(Exception):
        core.local_step_backward(z(N, 2).cpu(), z(N, 2), z(K, 2), z(K, 2, 2), z(K), z(K, 12), z(K, 10), S, z(N, K),
                                 z(N, K, S, 2), z(N, K), 1.0)


@pytest.mark.parametrize('dt', [torch.float64, torch.float32], ids=['f64', 'f32'])
@pytest.mark.parametrize('shape', [(100, 10, 2, 10), (274, 10, 6, 10), (33, 5, 11, 3)], ids=lambda s: 'N%dK%dD%dS%d' % s)
def test_backward_theta_record_gradient(shape, dt):
    """SMM variant: mu_k, L_k of the Student-t components are trained by gradient (experiments.py:154-174), so the
    reverse pass also returns d/d(theta record); checked against autograd through the oracle w.r.t. (W, m, cden)."""
    from oracle import backward as ob
    from vmp_for_svae_b200 import core
    from vmp_for_svae_b200.autograd import local_step_autograd, student_theta_record
    N, K, D, S = shape
    theta, phi_gmm, phi_enc, noise, gx, glr, greg = _inputs(N, K, D, S, seed=5 + N, student=True)
    W, m, cden, nu = ob.theta_consts_student(theta)
    lv = [t.clone().requires_grad_(True) for t in (W, m, cden)]
    x, log_r, reg = ob.forward(phi_enc[0], phi_enc[1], phi_gmm[0], phi_gmm[1], phi_gmm[2], lv[0], lv[1], lv[2], noise, nu=nu)
    ref = torch.autograd.grad((gx * x).sum() + (glr * log_r).sum() + greg * reg, lv)
    dev = lambda t: t.to(device=DEV, dtype=dt).contiguous()
    th = [dev(t) for t in theta]
    rec0 = core.theta_prepare_student(th)
    leaves = [dev(t).requires_grad_(True) for t in theta[1:3]]
    rec = student_theta_record(th[0], leaves[0], leaves[1], th[3])
    assert torch.allclose(rec, rec0, rtol=1e-10 if dt == torch.float64 else 1e-5, atol=1e-12 if dt == torch.float64 else 1e-6)
    rec_leaf = rec.detach().clone().requires_grad_(True)
    args = [dev(t) for t in (phi_enc[0], phi_enc[1], phi_gmm[0], phi_gmm[1], phi_gmm[2])]
    xg, lrg, regg, _ = local_step_autograd(*args, rec_leaf, S, den_mode=core.DEN_STUDENT, noise=dev(noise))
    (g_rec,) = torch.autograd.grad((dev(gx) * xg).sum() + (dev(glr) * lrg).sum() + greg * regg, [rec_leaf])
    g_rec = g_rec.cpu().double()
    tol = 1e-8 if dt == torch.float64 else 2e-3
    gW, gm, gc = g_rec[:, :D * D].reshape(K, D, D), g_rec[:, D * D:D * D + D], g_rec[:, D * D + D]
    for name, a, b in (('W', gW, torch.tril(ref[0])), ('m', gm, ref[1]), ('cden', gc, ref[2])):
        err = float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
        assert err < tol, (name, err)
    assert float(g_rec[:, D * D + D + 1:].abs().max()) == 0.0


@pytest.mark.parametrize('mode', ['standard', 'bernoulli'])
@pytest.mark.parametrize('dt', [torch.float64, torch.float32], ids=['f64', 'f32'])
def test_decoder_loglike_backward(mode, dt):
    """neg_rec term (vae.py:175-250) forward + reverse against autograd through the oracle restatement."""
    from oracle import svae_port as sp
    from vmp_for_svae_b200.autograd import decoder_loglike_autograd
    rs = np.random.RandomState(2)
    N, K, S, Do = 37, 5, 3, 45
    y = T(np.sign(rs.randn(N, Do))) if mode == 'bernoulli' else T(rs.randn(N, Do))
    means, out2 = T(rs.randn(N, K, S, Do)), T(rs.randn(N, K, S, Do))
    if mode == 'standard':
        out2 = torch.nn.functional.softplus(out2) + 0.05
    w = T(rs.dirichlet(np.ones(K), N))
    lv = [t.clone().requires_grad_(True) for t in (means, out2, w)]
    val = sp.expected_diagonal_gaussian_loglike(y, lv[0], lv[1], weights=lv[2]) if mode == 'standard' \
        else sp.expected_bernoulli_loglike(y, lv[1], r_nk=lv[2])
    ref = torch.autograd.grad(val, lv, allow_unused=True)
    dev = lambda t: t.to(device=DEV, dtype=dt).contiguous()
    lg = [dev(t).requires_grad_(True) for t in (means, out2, w)]
    got_val = decoder_loglike_autograd(dev(y), (lg[0], lg[1]), lg[2], mode)
    got = torch.autograd.grad(got_val * 1.5, lg, allow_unused=True)
    tol = 1e-10 if dt == torch.float64 else 2e-5
    assert abs(float(got_val) - float(val)) <= tol * abs(float(val))
    for a, b in zip(got, ref):
        if b is None:
            continue
        assert float((a.cpu().double() / 1.5 - b).abs().max() / b.abs().max()) < tol
