"""Test-time metrics with the reference's names and argument orders (losses.py), on CUDA tensors.

The O(N*K*S*Dobs) pass over the decoder outputs is ONE kernel (`vmp_decoder_metrics`, csrc/elbo_terms.cu) that yields the
per-(n,k) squared error and the per-(n,k) log-sum-exp over samples; what remains here is [N,K]-sized.
"""
import math

import numpy as np
import torch

from . import core


def weighted_mse(y_true, y_pred, r_nk_pred, name='mse'):
    """losses.py:9-40."""
    sq, _ = core.decoder_metrics(y_true, y_pred, y_pred, 0, want_lse=False)
    return (sq * r_nk_pred).sum(1).mean()


def _lse_nk(y, means, out2, log_weights, mask, mode):
    lw3 = log_weights if (log_weights is not None and log_weights.dim() == 3) else None
    _, lse = core.decoder_metrics(y, means, out2, mode, mask=mask, log_w_nks=lw3, want_sq=False)
    if log_weights is not None and log_weights.dim() == 2:
        lse = lse + log_weights
    return lse


def bernoulli_logprob(y_true_bin, logits, log_weights=None, missing_data_mask=None, name='bernoulli_logprob'):
    """losses.py:43-80 ([N,K,S,D] logits with log_weights; [N,S,D] without).  The reference subtracts S, not log S,
    after the log-sum-exp over samples (75-77); kept."""
    if log_weights is None:
        logits = logits.unsqueeze(1)
    S = logits.shape[2]
    lse = _lse_nk(y_true_bin, logits, logits, log_weights, missing_data_mask, 1)
    return (torch.logsumexp(lse, dim=1) - float(S)).mean()


def diagonal_gaussian_logprob(y_true, mean, var, log_weights, mask=None, name='gauss_logprob'):
    """losses.py:83-144."""
    S = mean.shape[2]
    lse = _lse_nk(y_true, mean, var, log_weights, mask, 0)
    return torch.logsumexp(lse - math.log(S), dim=1).mean()


def imputation_mse(y_true, y_pred, r_nk_pred, missing_data_mask, name='imp_mse'):
    """losses.py:147-170."""
    sq, _ = core.decoder_metrics(y_true, y_pred, y_pred, 0, mask=missing_data_mask, want_lse=False)
    return (sq * r_nk_pred).sum() / y_true.shape[0]


def generate_missing_data_mask(y, noise_ratio=0.3, mask_type='random', seed=0, name='make_mask'):
    """losses.py:242-274 -> bool[N,D] on y's device (constant for a run)."""
    N, D = y.shape
    mask = np.zeros(N * D, dtype=bool)
    if mask_type == 'random':
        idx = np.random.RandomState(seed).choice(np.arange(N * D), size=int(N * D * noise_ratio), replace=False)
        mask[idx] = True
    else:
        side = int(round(math.sqrt(D)))
        assert side * side == D
        half = side // 2
        mask = mask.reshape(N, side, side)
        if mask_type == 'quarter':
            mask[:, half:side, :half] = True
        elif mask_type == 'lower_half':
            mask[:, half:side, :side] = True
        elif mask_type == 'left_half':
            mask[:, :side, :half] = True
        else:
            raise NotImplementedError("The mask type '%s' does not exist." % mask_type)
    return torch.as_tensor(mask.reshape(N, D), device=y.device)


def perturb_data(y, missing_data_mask, seed, decoder_type='standard', name='perturb_data', *, noise=None):
    """losses.py:277-310 : missing entries replaced by N(0,1) noise ('standard') or +-1 coin flips ('bernoulli')."""
    if noise is None:
        g = torch.Generator(device=y.device).manual_seed(int(seed))
        if decoder_type == 'standard':
            noise = torch.randn(y.shape, generator=g, device=y.device, dtype=y.dtype)
        elif decoder_type == 'bernoulli':
            noise = torch.randint(0, 2, y.shape, generator=g, device=y.device).to(y.dtype) * 2.0 - 1.0
        else:
            raise NotImplementedError
    m = missing_data_mask.to(y.dtype)
    return (1.0 - m) * y + m * noise


def imputation_losses(y_true, missing_data_mask, imputation_method, nb_samples_pert=100, nb_samples_rec=100, seed=0,
                      decoder_type='standard', name='imputation_losses', *, noises=None):
    """losses.py:173-239.  `imputation_method(y_perturbed) -> (means, vars|logits, log_r)`.  As in the reference,
    perturb_data is called WITHOUT decoder_type (line 213), so the fill-in noise is Gaussian for both decoders.
    `noises[p]` injects the draws (tests); otherwise perturbation p uses seed + p."""
    y01 = torch.where(y_true == -1, torch.zeros_like(y_true), torch.ones_like(y_true)) if decoder_type == 'bernoulli' else y_true
    mode = 1 if decoder_type == 'bernoulli' else 0
    mse, lses = 0.0, []
    S_tot = 0
    for p in range(nb_samples_pert):
        y_pert = perturb_data(y_true, missing_data_mask, seed + p, noise=None if noises is None else noises[p])
        means, out2, log_r = imputation_method(y_pert)
        sq, lse = core.decoder_metrics(y_true, means, out2, mode, target=y01, mask=missing_data_mask)
        mse = mse + (sq * torch.exp(log_r)).sum() / y_true.shape[0]
        lses.append(lse + log_r)
        S_tot += means.shape[2]
    lse = torch.logsumexp(torch.stack(lses, 0), dim=0)                  # concat over the sample axis (225-227)
    if decoder_type == 'bernoulli':
        ll = (torch.logsumexp(lse, dim=1) - float(S_tot)).mean()
    else:
        ll = torch.logsumexp(lse - math.log(S_tot), dim=1).mean()
    return mse / nb_samples_pert, ll


def purity(r_nk, labels, eps=1e-10, name='purity'):
    """losses.py:313-349 ; labels one-hot [N,C] -> (entropy, purity)."""
    N = r_nk.shape[0]
    N_kc = r_nk.t() @ labels.to(r_nk.dtype)
    N_k = r_nk.sum(0)
    p_kc = N_kc / (N_k + eps).unsqueeze(1)
    ent_k = -(p_kc * torch.log(p_kc + eps)).sum(1)
    return (N_k / N * ent_k).sum(), (N_k / N * p_kc.max(1).values).sum()
