"""End-to-end training through the CUDA forward + reverse kernels: the behaviour of the reference's
experiments.py on the pinwheel data — the ELBO improves and the reconstruction error falls."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('method', ['svae-cvi', 'svae-cvi-smm'])
def test_pinwheel_training_improves(method):
    from vmp_for_svae_b200 import experiments as ex
    cfg = dict(dataset='pinwheel', method=method, lr=0.01, lrcvi=0.1, decay_rate=1.0, K=10, L=2, U=40, DoF=5, seed=0)
    tr, hist = ex.run_experiment(cfg, nb_iters=400, size_minibatch=100, measurement_freq=100, verbose=False,
                                 nb_samples_te=20)
    first, last = hist[0], hist[-1]
    assert all(h['bad_pivots'] == 0 for h in hist)
    assert all(torch.isfinite(torch.tensor([h['neg_elbo_normed'], h['mse']])).all() for h in hist)
    assert last['neg_elbo_normed'] < first['neg_elbo_normed'] - 1.0, (first, last)
    assert last['mse'] < 0.5 * first['mse'], (first, last)


def test_one_step_matches_oracle_graph():
    """One training iteration's gradients on every trainable tensor (encoder, decoder, phi_gmm) against the same
    graph built from the oracle's differentiable forward with torch autograd, fp32 nets in fp64 copies."""
    import copy
    import numpy as np
    from oracle import backward as ob, svae_port as sp
    from vmp_for_svae_b200 import core, experiments as ex
    from vmp_for_svae_b200.autograd import decoder_loglike_autograd, local_step_autograd
    cfg = dict(dataset='pinwheel', method='svae-cvi', lr=0.01, lrcvi=0.1, K=6, L=3, U=20, seed=1)
    tr = ex.SVAETrainer(cfg, obs_dim=4, device='cuda:0', nb_samples=3, stddev_init_nn=0.3)
    for m in (tr.encoder, tr.decoder):
        m.double()
    tr.phi_gmm = [torch.nn.Parameter(p.detach().double()) for p in tr.phi_gmm]
    tr.theta = [t.double() for t in tr.theta]
    y = torch.as_tensor(np.random.RandomState(0).randn(50, 4) * 2.0, dtype=torch.float64, device='cuda:0')
    N, K, D, S = 50, 6, 3, 3
    noise, _ = core.fill_noise(N, K, D, S, 99, torch.float64, 'cuda:0', want_u=False)
    eta1, eta2d = tr.encoder(y)
    x_k, log_r, reg, _ = local_step_autograd(eta1, eta2d, *tr.phi_gmm, core.theta_prepare_gauss(tr.theta), S, noise=noise)
    elbo = decoder_loglike_autograd(y, tr.decoder(x_k), torch.exp(log_r), 'standard') - reg
    params = list(tr.encoder.parameters()) + list(tr.decoder.parameters()) + tr.phi_gmm
    got = torch.autograd.grad(-elbo, params)
    # the oracle graph on CPU
    enc, dec = copy.deepcopy(tr.encoder).cpu(), copy.deepcopy(tr.decoder).cpu()
    phi = [p.detach().cpu().clone().requires_grad_(True) for p in tr.phi_gmm]
    W, m, cden = ob.theta_consts_gauss([t.cpu() for t in tr.theta])
    e1, e2 = enc(y.cpu())
    xo, lro, rego = ob.forward(e1, e2, phi[0], phi[1], phi[2], W, m, cden, noise.cpu())
    means, vars_ = dec(xo)
    elbo_o = sp.expected_diagonal_gaussian_loglike(y.cpu(), means, vars_, weights=torch.exp(lro)) - rego
    ref = torch.autograd.grad(-elbo_o, list(enc.parameters()) + list(dec.parameters()) + phi)
    assert abs(float(elbo) - float(elbo_o)) < 1e-8 * abs(float(elbo_o))
    for a, b in zip(got, ref):
        assert float((a.cpu() - b).abs().max()) <= 1e-7 * max(float(b.abs().max()), 1e-6)


@pytest.mark.parametrize('smm', [False, True], ids=['gmm', 'smm'])
def test_dropin_surface_is_differentiable(smm):
    """The reference-shaped calls svae.inference + svae.compute_elbo(_smm) + (-elbo).backward() (experiments.py:208-232)
    give the oracle graph's gradients on encoder, decoder, phi_gmm (and mu_k, L_k of the SMM variant)."""
    import copy
    import numpy as np
    from oracle import backward as ob, svae_port as sp
    from vmp_for_svae_b200 import core, experiments as ex
    from vmp_for_svae_b200.models import svae
    dev = 'cuda:0'
    N, K, D, S, Do = 40, 5, 3, 2, 4
    cfg = dict(dataset='pinwheel', method='svae-cvi-smm' if smm else 'svae-cvi', lr=0.01, lrcvi=0.1, K=K, L=D, U=16,
               DoF=5, seed=2)
    tr = ex.SVAETrainer(cfg, obs_dim=Do, device=dev, nb_samples=S, stddev_init_nn=0.3)
    tr.encoder.double(); tr.decoder.double()
    phi_gmm = [torch.nn.Parameter(p.detach().double()) for p in tr.phi_gmm]
    if smm:
        theta = (tr.alpha.double(), torch.nn.Parameter(tr.mu_k.detach().double()),
                 torch.nn.Parameter(tr.L_k.detach().double()), tr.dof.double())
        extra = [theta[1], theta[2]]
    else:
        theta, extra = [t.double() for t in tr.theta], []
    y = torch.as_tensor(np.random.RandomState(1).randn(N, Do) * 2.0, dtype=torch.float64, device=dev)
    noise, _ = core.fill_noise(N, K, D, S, 5, torch.float64, dev, want_u=False)
    y_rec, _, x_k, x_samples, log_z, _, phi_tilde = svae.inference(y, phi_gmm, tr.encoder, tr.decoder, nb_samples=S,
                                                                   seed=0, noise=noise)
    fn = svae.compute_elbo_smm if smm else svae.compute_elbo
    elbo, details = fn(y, y_rec, theta, phi_tilde, x_k, log_z, 'standard')
    params = list(tr.encoder.parameters()) + list(tr.decoder.parameters()) + phi_gmm + extra
    got = torch.autograd.grad(-elbo, params)
    assert x_samples.shape == (N, D)
    # oracle graph
    enc, dec = copy.deepcopy(tr.encoder).cpu(), copy.deepcopy(tr.decoder).cpu()
    phi = [p.detach().cpu().clone().requires_grad_(True) for p in phi_gmm]
    if smm:
        th = [t.detach().cpu().clone() for t in theta]
        th[1].requires_grad_(True); th[2].requires_grad_(True)
        W, m, cden, nu = ob.theta_consts_student(th)
        oextra = [th[1], th[2]]
    else:
        (W, m, cden), nu, oextra = ob.theta_consts_gauss([t.cpu() for t in theta]), None, []
    e1, e2 = enc(y.cpu())
    xo, lro, rego = ob.forward(e1, e2, phi[0], phi[1], phi[2], W, m, cden, noise.cpu(), nu=nu)
    means, vars_ = dec(xo)
    elbo_o = sp.expected_diagonal_gaussian_loglike(y.cpu(), means, vars_, weights=torch.exp(lro)) - rego
    ref = torch.autograd.grad(-elbo_o, list(enc.parameters()) + list(dec.parameters()) + phi + oextra)
    assert abs(float(elbo.detach()) - float(elbo_o.detach())) < 1e-8 * abs(float(elbo_o.detach()))
    for a, b in zip(got, ref):
        assert float((a.cpu() - b).abs().max()) <= 1e-7 * max(float(b.abs().max()), 1e-6)


@pytest.mark.parametrize('method', ['svae-cvi', 'svae-cvi-smm'])
def test_graphed_trainer_trains_and_is_faster(method):
    """The CUDA-graph form of the training iteration: same behaviour (ELBO improves, MSE falls) at a fraction of the
    per-iteration time of the eager trainer."""
    import time
    from vmp_for_svae_b200 import experiments as ex
    cfg = dict(dataset='pinwheel', method=method, lr=0.01, lrcvi=0.1, decay_rate=0.95, K=10, L=2, U=40, DoF=5, seed=0)
    X_tr, _, X_te, l_te = ex.make_dataset('pinwheel')
    dev = torch.device('cuda', 0)
    y_tr = torch.as_tensor(X_tr, dtype=torch.float32, device=dev)
    y_te = torch.as_tensor(X_te, dtype=torch.float32, device=dev)
    torch.manual_seed(0)
    tr = ex.GraphedSVAETrainer(cfg, y_tr, 100, device=dev).capture()
    first = tr.evaluate(y_te, torch.as_tensor(l_te, device=dev), nb_samples=20)
    e0 = float(tr.train_step()[0])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(600):
        out = tr.train_step()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 600 * 1e3
    last = tr.evaluate(y_te, torch.as_tensor(l_te, device=dev), nb_samples=20)
    assert float(out[3]) == 0 and torch.isfinite(out).all()
    assert float(out[0]) > e0 + 100.0, (e0, float(out[0]))
    assert last['mse'] < 0.5 * first['mse'], (first, last)
    assert abs(float(tr.rho) - 0.1 * 0.95 ** (tr.global_step / 1000.0)) < 1e-9
    print('graphed training iteration: %.3f ms' % ms)
    assert ms < 2.0, ms            # eager iteration: 2.2-2.6 ms; measured replay: 0.52 ms (GMM), 0.65 ms (SMM)


@pytest.mark.parametrize('method,L,U', [('svae-cvi', 3, 20), ('svae-cvi-smm', 2, 16), ('svae-cvi', 32, 24)],
                         ids=['gmm-L3', 'smm-L2', 'gmm-L32'])
def test_streamed_iteration_equals_whole_batch(method, L, U):
    """SVAETrainer.train_step_streamed (x_k_samples produced, decoded and back-propagated tile by tile: the [N,K,S,D] tensor of
    svae.py:511 never exists for the whole batch) == train_step on the whole batch: same ELBO, same parameter gradients, same
    CVI update.  L = 32 also exercises the block-cooperative reverse kernel inside a training iteration (C4-shaped latent)."""
    import copy
    import numpy as np
    from vmp_for_svae_b200 import experiments as ex
    cfg = dict(dataset='pinwheel', method=method, lr=0.01, lrcvi=0.1, K=7, L=L, U=U, DoF=5, seed=2)
    a = ex.SVAETrainer(cfg, obs_dim=5, device='cuda:0', nb_samples=2, stddev_init_nn=0.2)
    b = copy.deepcopy(a)
    y = torch.as_tensor(np.random.RandomState(1).randn(83, 5) * 1.5, dtype=torch.float32, device='cuda:0')
    grads = []
    for tr, step in ((a, lambda t: t.train_step(y)), (b, lambda t: t.train_step_streamed(y, 17))):
        caught = {}
        orig = tr.opt.step
        tr.opt.step = lambda tr=tr, caught=caught, orig=orig: caught.update(
            g=[p.grad.detach().clone() for grp in tr.opt.param_groups for p in grp['params']]) or orig()
        out = step(tr)
        grads.append((out, caught['g'], [t.clone() for t in (tr.theta if not tr.smm else [tr.alpha])]))
    (oa, ga, ta), (ob_, gb, tb) = grads
    assert float(oa['bad_pivots']) == 0 and float(ob_['bad_pivots']) == 0
    assert abs(float(oa['elbo']) - float(ob_['elbo'])) <= 2e-5 * abs(float(oa['elbo']))
    for x, z in zip(ga, gb):
        assert float((x - z).abs().max()) <= 2e-4 * max(float(x.abs().max()), 1e-6)
    for x, z in zip(ta, tb):
        torch.testing.assert_close(z, x, rtol=2e-5, atol=1e-5)
