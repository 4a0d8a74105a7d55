"""distributions/student_t.py of the reference.

In the SMM-SVAE hot path the Student-t denominator is evaluated inside the fused local-step kernel
(VMP_DEN_STUDENT).  The general functions below (arbitrary y) use the batched Cholesky kernel for the scale
matrices and K-sized tensor algebra for the rest."""
import math

import torch

from .. import core


def _logprob_full_scale(y, mu, sigma, v, name='student_t_logprob'):
    """student_t.py:7-39 : y[N,K,S,D], mu[K,D], sigma[K,D,D], v[K] -> [N,K,S]."""
    N, K, S, D = y.shape
    assert tuple(mu.shape) == (K, D)
    assert tuple(sigma.shape) == (K, D, D)
    assert tuple(v.shape) == (K,)
    inv, logdet = core.spd_inverse(sigma)
    err = y - mu.unsqueeze(0).unsqueeze(2)
    maha = torch.einsum('nksd,kde,nkse->nks', err, inv, err)
    vv = v.unsqueeze(0).unsqueeze(2)
    logprob = torch.lgamma(0.5 * (vv + D)) - torch.lgamma(0.5 * vv)
    logprob = logprob - 0.5 * D * torch.log(math.pi * vv)
    logprob = logprob - 0.5 * logdet.unsqueeze(0).unsqueeze(2)
    logprob = logprob - 0.5 * (vv + D) * torch.log1p(maha / vv)
    return logprob


def logprob_smm_mixture(y, mu, sigma, v, log_pi, name='student_t_logprob'):
    """student_t.py:42-56 : y[N,D] -> [N,K]."""
    N, D = y.shape
    K, D_ = mu.shape
    assert D_ == D
    yy = y.unsqueeze(1).unsqueeze(2).expand(N, K, 1, D)
    return _logprob_full_scale(yy, mu, sigma, v).reshape(N, K) + log_pi.unsqueeze(0)


def log_probability_per_samp(y, mu, sigma, v, name='student_t_logprob_per_samp'):
    """student_t.py:59-61."""
    return _logprob_full_scale(y, mu, sigma, v)
