"""distributions/gaussian.py of the reference.

The conversions are K-sized; the two log-densities take dense natural parameters eta2[N,K,D,D] and run in
elbo_terms.cu (one Cholesky per (n,k)).  The SVAE hot path never builds eta2[N,K,D,D]: models.svae uses the fused
local-step kernel instead."""
import torch

from .. import core


def standard_to_natural(mu, sigma, name='gauss_to_nat'):
    """gaussian.py:11-19 : eta2 = -1/2 inv(sigma), eta1 = -2 eta2 mu."""
    inv, _ = core.spd_inverse(sigma, want_logdet=False)
    eta_2 = -0.5 * inv
    eta_1 = (-2 * eta_2 @ mu.unsqueeze(-1)).reshape(mu.shape)
    return eta_1, eta_2


def natural_to_standard(eta1, eta2, name='gauss_to_stndrd'):
    """gaussian.py:22-27 : sigma = inv(-2 eta2), mu = sigma eta1."""
    sigma, _ = core.spd_inverse(-2 * eta2, want_logdet=False)
    mu = (sigma @ eta1.unsqueeze(-1)).reshape(eta1.shape)
    return mu, sigma


def log_probability_nat(x, eta1, eta2, weights=None):
    """gaussian.py:30-71 : x[N,D], eta1[N,K,D], eta2[N,K,D,D], weights[K] -> [N,K], normalised over K."""
    if eta1.dim() != 3:
        raise AssertionError("eta1 must be of shape (N,K,D). Its shape is %s." % str(tuple(eta1.shape)))
    log_w = torch.log(weights) if weights is not None else None
    return core.gaussian_logprob_nat(x, eta1, eta2, log_w, per_samp=False)


def log_probability_nat_per_samp(x_samps, eta1, eta2):
    """gaussian.py:74-105 : x[N,K,S,D], eta1[N,K,D], eta2[N,K,D,D] -> [N,K,S]."""
    N, K, S, D = x_samps.shape
    assert tuple(eta1.shape) == (N, K, D)
    assert tuple(eta2.shape) == (N, K, D, D)
    return core.gaussian_logprob_nat(x_samps, eta1, eta2, None, per_samp=True)
