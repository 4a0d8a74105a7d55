"""oracle/losses_port.py against the fixtures produced by the reference's own losses.py (tests/golden/losses.npz)."""
import numpy as np
import torch

from conftest import load_golden, T, losses_inputs, imputation_stub
from oracle import losses_port as lp


def _close(a, b, tol=1e-10):
    assert abs(float(a) - float(b)) <= tol * max(1.0, abs(float(b))), (float(a), float(b))


def test_losses_port_vs_reference_golden():
    g = load_golden('losses')
    N, K, S, D, P = (int(g[k]) for k in 'NKSDP')
    y, pred, var, logits, r, labels = (T(a) if a.dtype.kind == 'f' else a for a in losses_inputs(int(g['seed']), N, K, S, D))
    yb = torch.sign(y)
    mask = lp.generate_missing_data_mask(N, D, 0.3, seed=0)
    assert np.array_equal(mask.numpy(), g['mask'])
    _close(lp.weighted_mse(y, pred, r), g['weighted_mse'])
    _close(lp.diagonal_gaussian_logprob(y, pred, var, torch.log(r)), g['gauss_logprob'])
    lw3 = torch.log(r)[:, :, None] + 0.1 * T(np.random.RandomState(3).randn(N, K, S))
    _close(lp.diagonal_gaussian_logprob(y, pred, var, lw3), g['gauss_logprob_nks'])
    _close(lp.diagonal_gaussian_logprob(y, pred, var, torch.log(r), mask=mask), g['gauss_logprob_mask'])
    _close(lp.bernoulli_logprob(yb, logits, torch.log(r)), g['bernoulli_logprob'])
    _close(lp.bernoulli_logprob(yb, logits, torch.log(r), mask), g['bernoulli_logprob_mask'])
    _close(lp.imputation_mse(y, pred, r, mask), g['imputation_mse'])
    ent, pur = lp.purity(r, T(np.eye(4)[labels]))
    _close(ent, g['entropy']); _close(pur, g['purity'])
    for dt, yy in (('standard', y), ('bernoulli', yb)):
        stub = imputation_stub(N, K, S, D)
        mse, ll = lp.imputation_losses(yy, mask, lambda yp: tuple(T(t) for t in stub(yp.numpy())), T(g['imp_noise']), dt)
        _close(mse, g['imp_mse_' + dt]); _close(ll, g['imp_ll_' + dt])
