"""vmp_for_svae_b200.losses (CUDA `vmp_decoder_metrics` + [N,K]-sized tails) against the reference's own losses.py
outputs (tests/golden/losses.npz) and against the oracle restatement at a second, ragged shape."""
import numpy as np
import pytest
import torch

from conftest import load_golden, T, losses_inputs, imputation_stub

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _close(a, b, tol):
    assert abs(float(a) - float(b)) <= tol * max(1.0, abs(float(b))), (float(a), float(b))


@pytest.mark.parametrize('dt', [torch.float64, torch.float32], ids=['f64', 'f32'])
def test_losses_vs_reference_golden(dt):
    from vmp_for_svae_b200 import losses
    g = load_golden('losses')
    tol = 1e-10 if dt == torch.float64 else 2e-5
    N, K, S, D, P = (int(g[k]) for k in 'NKSDP')
    y, pred, var, logits, r, labels = losses_inputs(int(g['seed']), N, K, S, D)
    dev = lambda a: torch.as_tensor(a).to(device=DEV, dtype=dt).contiguous()
    y, pred, var, logits, r = (dev(a) for a in (y, pred, var, logits, r))
    yb = torch.sign(y)
    mask = losses.generate_missing_data_mask(y, 0.3, seed=0)
    assert np.array_equal(mask.cpu().numpy(), g['mask'])
    _close(losses.weighted_mse(y, pred, r), g['weighted_mse'], tol)
    _close(losses.diagonal_gaussian_logprob(y, pred, var, torch.log(r)), g['gauss_logprob'], tol)
    lw3 = torch.log(r)[:, :, None] + 0.1 * dev(np.random.RandomState(3).randn(N, K, S))
    _close(losses.diagonal_gaussian_logprob(y, pred, var, lw3.contiguous()), g['gauss_logprob_nks'], tol)
    _close(losses.diagonal_gaussian_logprob(y, pred, var, torch.log(r), mask=mask), g['gauss_logprob_mask'], tol)
    _close(losses.bernoulli_logprob(yb, logits, torch.log(r)), g['bernoulli_logprob'], tol)
    _close(losses.bernoulli_logprob(yb, logits, torch.log(r), mask), g['bernoulli_logprob_mask'], tol)
    _close(losses.imputation_mse(y, pred, r, mask), g['imputation_mse'], tol)
    ent, pur = losses.purity(r, dev(np.eye(4)[labels]))
    _close(ent, g['entropy'], tol); _close(pur, g['purity'], tol)
    for kind, yy in (('standard', y), ('bernoulli', yb)):
        stub = imputation_stub(N, K, S, D)
        mse, ll = losses.imputation_losses(yy, mask, lambda yp: tuple(dev(t) for t in stub(yp.cpu().double().numpy())),
                                           P, S, decoder_type=kind, noises=dev(g['imp_noise']))
        _close(mse, g['imp_mse_' + kind], 10 * tol); _close(ll, g['imp_ll_' + kind], tol)


def test_losses_vs_oracle_ragged_shape():
    from oracle import losses_port as lp
    from vmp_for_svae_b200 import losses
    N, K, S, D = 61, 7, 3, 45                                 # Dobs not a multiple of the warp width
    y, pred, var, logits, r, labels = losses_inputs(5, N, K, S, D)
    c = [T(a) for a in (y, pred, var, logits, r)]
    d = [t.to(DEV) for t in c]
    mask = lp.generate_missing_data_mask(N, D, 0.2, seed=4)
    _close(losses.weighted_mse(d[0], d[1], d[4]), lp.weighted_mse(c[0], c[1], c[4]), 1e-10)
    _close(losses.diagonal_gaussian_logprob(d[0], d[1], d[2], torch.log(d[4]), mask=mask.to(DEV)),
           lp.diagonal_gaussian_logprob(c[0], c[1], c[2], torch.log(c[4]), mask=mask), 1e-10)
    _close(losses.bernoulli_logprob(torch.sign(d[0]), d[3][:, 0], None),
           lp.bernoulli_logprob(torch.sign(c[0]), c[3][:, 0], None), 1e-10)
    _close(losses.imputation_mse(d[0], d[1], d[4], mask.to(DEV)), lp.imputation_mse(c[0], c[1], c[4], mask), 1e-10)
