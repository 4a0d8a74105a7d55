"""ORACLE (test infrastructure): torch-CPU restatement of the reference's test-time metrics (losses.py), pinned
against fixtures produced by executing the reference's own losses.py (tests/golden/losses.npz, make_golden.py)."""
import math

import numpy as np
import torch


def weighted_mse(y_true, y_pred, r_nk):
    """losses.py:9-40 : mean_n sum_k r_nk mean_s |y_n - yhat_nks|^2."""
    se = ((y_true[:, None, None, :] - y_pred) ** 2).sum(3).mean(2)
    return (se * r_nk).sum(1).mean()


def bernoulli_logprob(y_bin, logits, log_weights=None, missing_data_mask=None):
    """losses.py:43-80.  NB the reference subtracts S (not log S) after the log-sum-exp over samples (line 75-77)."""
    S = logits.shape[-2]
    yb = y_bin[:, None, :] if log_weights is None else y_bin[:, None, None, :]
    pix = -torch.log(1.0 + torch.exp(-logits * yb))
    if missing_data_mask is not None:
        m = missing_data_mask.to(pix.dtype)[:, None, :]
        pix = pix * (m if log_weights is None else m[:, None])
    lp = pix.sum(-1)
    if log_weights is not None:
        lw = log_weights if log_weights.dim() == 3 else log_weights[:, :, None]
        lp = torch.logsumexp(lp + lw, dim=1)
    return (torch.logsumexp(lp, dim=-1) - float(S)).mean()


def diagonal_gaussian_logprob(y_true, mean, var, log_weights, mask=None):
    """losses.py:83-144 : mean_n log sum_k [ 1/S sum_s w_nk(s) N(y_n | mean_nks, var_nks) ] (masked dims dropped)."""
    S = mean.shape[2]
    lp = -0.5 * ((y_true[:, None, None, :] - mean) ** 2 / var + torch.log(var) + math.log(2 * math.pi))
    if mask is not None:
        lp = lp * mask.to(lp.dtype)[:, None, None, :]
    lw = log_weights if log_weights.dim() == 3 else log_weights[:, :, None]
    lp = lp.sum(3) + lw
    lp_k = torch.logsumexp(lp, dim=2) - math.log(S)
    return torch.logsumexp(lp_k, dim=1).mean()


def imputation_mse(y_true, y_pred, r_nk, missing_data_mask):
    """losses.py:147-170."""
    m = missing_data_mask.to(y_true.dtype)
    se = (((y_true * m)[:, None, None, :] - y_pred * m[:, None, None, :]) ** 2).mean(2)
    return (se * r_nk[:, :, None]).sum() / y_true.shape[0]


def generate_missing_data_mask(N, D, noise_ratio=0.3, seed=0):
    """losses.py:242-258 ('random' mask)."""
    mask = np.zeros(N * D, dtype=bool)
    idx = np.random.RandomState(seed).choice(np.arange(N * D), size=int(N * D * noise_ratio), replace=False)
    mask[idx] = True
    return torch.as_tensor(mask.reshape(N, D))


def perturb_data(y, missing_data_mask, noise):
    """losses.py:277-310 : missing entries <- noise (imputation_losses never forwards decoder_type, so the noise is
    N(0,1) for both decoder types, line 213), observed entries kept."""
    m = missing_data_mask.to(y.dtype)
    return (1.0 - m) * y + m * noise


def imputation_losses(y_true, missing_data_mask, imputation_method, noises, decoder_type='standard'):
    """losses.py:173-239 ; `noises[p]` is the N(0,1) draw of perturbation p (injected)."""
    y01 = torch.where(y_true == -1, torch.zeros_like(y_true), torch.ones_like(y_true)) if decoder_type == 'bernoulli' else y_true
    m4 = missing_data_mask.to(y_true.dtype)[:, None, None, :]
    mse, means, outs, lws = 0.0, [], [], []
    for p in range(len(noises)):
        mean, out2, log_r = imputation_method(perturb_data(y_true, missing_data_mask, noises[p]))
        mse = mse + imputation_mse(y01, mean * m4, torch.exp(log_r), missing_data_mask)
        means.append(mean); outs.append(out2); lws.append(log_r[:, :, None].expand(-1, -1, mean.shape[2]))
    means, outs, lws = torch.cat(means, 2), torch.cat(outs, 2), torch.cat(lws, 2)
    if decoder_type == 'bernoulli':
        ll = bernoulli_logprob(y_true, outs, lws, missing_data_mask)
    else:
        ll = diagonal_gaussian_logprob(y_true, means, outs, lws, mask=missing_data_mask)
    return mse / len(noises), ll


def purity(r_nk, labels_1h, eps=1e-10):
    """losses.py:313-349."""
    N = r_nk.shape[0]
    N_kc = torch.einsum('nk,nc->kc', r_nk, labels_1h)
    N_k = r_nk.sum(0)
    p_kc = N_kc / (N_k + eps)[:, None]
    ent_k = -(p_kc * torch.log(p_kc + eps)).sum(1)
    return (N_k / N * ent_k).sum(), (N_k / N * p_kc.max(1).values).sum()
