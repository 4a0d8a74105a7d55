"""Drop-in surface of the reference's models/gmm.py (Bayesian GMM, Bishop PRML 10.2) on B200.

N-sized work runs in libvmp_svae.so: the responsibility-weighted moments (suffstats.cu, one pass over the data for
N_k, x_k and S_k together), the standard-parameter M-step (vmp_mixture_mstep) and the E-step (mixtures.cu).
The small `update_*` helpers keep the reference's names; those that only combine K-sized quantities are K-sized
tensor expressions.
"""
import torch

from .. import core
from ..lazy import lazy_log
from ..distributions import dirichlet, niw
from . import svae


def _moments(x, w):
    """(sum_n w, sum_n w x, sum_n w x x^T) from the stats kernel, in the working dtype."""
    N, D = x.shape
    st = core.suffstats(x, w, r_is_log=False)
    K = st.shape[0]
    return (st[:, 1].to(x.dtype), st[:, 2:2 + D].to(x.dtype), st[:, 2 + D:].reshape(K, D, D).to(x.dtype))


def update_Nk(r_nk):
    """gmm.py:25-27 (Bishop 10.51)."""
    N, K = r_nk.shape
    st = core.suffstats(torch.zeros(N, 1, dtype=r_nk.dtype, device=r_nk.device), r_nk)
    return st[:, 0].to(r_nk.dtype)


def update_xk(x, r_nk, N_k):
    """gmm.py:30-36 (Bishop 10.52), NaN -> unnormalised."""
    _, s1, _ = _moments(x, r_nk)
    normed = s1 / N_k.unsqueeze(1)
    return torch.where(torch.isnan(normed), s1, normed)


def update_Sk(x, r_nk, N_k, x_k):
    """gmm.py:39-46 (Bishop 10.53), NaN -> unnormalised."""
    s0, s1, s2 = _moments(x, r_nk)
    S = s2 - x_k.unsqueeze(2) * s1.unsqueeze(1) - s1.unsqueeze(2) * x_k.unsqueeze(1) \
        + s0.reshape(-1, 1, 1) * x_k.unsqueeze(2) * x_k.unsqueeze(1)
    normed = S / N_k.reshape(-1, 1, 1)
    return torch.where(torch.isnan(normed), S, normed)


def update_alphak(alpha_0, N_k):
    """gmm.py:49-51 (Bishop 10.58)."""
    return alpha_0 + N_k


def update_betak(beta_0, N_k):
    """gmm.py:54-56 (Bishop 10.60)."""
    return beta_0 + N_k


def update_mk(beta_0, m_0, N_k, x_k, beta_k):
    """gmm.py:59-67 (Bishop 10.61)."""
    if beta_0.dim() == 1:
        beta_0 = beta_0.reshape(-1, 1)
    return (beta_0 * m_0 + N_k.unsqueeze(1) * x_k) / beta_k.unsqueeze(1)


def update_Ck(C_0, x_k, N_k, m_0, beta_0, beta_k, S_k):
    """gmm.py:70-76 (Bishop 10.62)."""
    Q0 = x_k - m_0
    return C_0 + N_k.reshape(-1, 1, 1) * S_k + (beta_0 * N_k / beta_k).reshape(-1, 1, 1) * Q0.unsqueeze(2) * Q0.unsqueeze(1)


def update_vk(v_0, N_k):
    """gmm.py:79-81 (Bishop 10.63 with the reference's +1)."""
    return v_0 + N_k + 1


def compute_log_pi(alpha_k):
    """gmm.py:134-138 (Bishop 10.66)."""
    return torch.special.digamma(alpha_k) - torch.special.digamma(alpha_k.sum())


def e_step(x, alpha_k, beta_k, m_k, P_k, v_k, name='e_step'):
    """gmm.py:154-174 -> (r_nk[N,K], exp(E log pi)[K])."""
    r, _, pi = core.mixture_estep(x, alpha_k, beta_k, m_k, P_k, v_k)
    return r, pi


def e_step_missing_data(x, alpha_k, beta_k, m_k, P_k, v_k, missing_data_mask, name='e_step_imp'):
    """gmm.py:177-198 (residuals of missing entries are zeroed, gmm.py:106-108)."""
    r, _, pi = core.mixture_estep(x, alpha_k, beta_k, m_k, P_k, v_k, missing_mask=missing_data_mask)
    return r, pi


def m_step(x, r_nk, alpha_0, beta_0, m_0, C_0, v_0, name='m_step'):
    """gmm.py:201-227 -> (alpha_k, beta_k, m_k, C_k, v_k, x_k, S_k)."""
    N, D = x.shape
    stats = core.suffstats(x, r_nk, r_is_log=False)
    return core.mixture_mstep(stats, D, False, alpha_0, beta_0, m_0, C_0, v_0)


_PRIOR_CACHE = {}


def _prior_standard(K, D, seed, dtype, device):
    """The sweep's Dirichlet+NIW prior in standard parameters (gmm.py: init_mm_params(...) -> niw.natural_to_standard,
    dirichlet.natural_to_standard); constant for a given (K, D), so it is built once instead of every sweep."""
    key = (K, D, int(seed), dtype, str(device))
    if key not in _PRIOR_CACHE:
        alpha, A, b, beta, v_hat = svae.init_mm_params(K, D, alpha_scale=0.05 / K, beta_scale=0.5, m_scale=0,
                                                       C_scale=D + 0.5, v_init=D + 0.5, seed=seed, device=device,
                                                       dtype=dtype)
        beta_0, m_0, C_0, v_0 = niw.natural_to_standard(A, b, beta, v_hat)
        alpha_0 = dirichlet.natural_to_standard(alpha)
        _PRIOR_CACHE[key] = tuple(t.contiguous() for t in (alpha_0, beta_0, m_0, C_0, v_0))
    return _PRIOR_CACHE[key]



def inference(x, K, seed, name='inference', *, r_nk=None, dtype=None):
    """gmm.py:230-269 : one VB-EM sweep.  The reference keeps r_nk in a tf.Variable initialised from Dirichlet(1);
    here the state tensor is passed in (`r_nk=`, updated IN PLACE) or created on first use.
    Returns (r_nk (state, updated), log_r_nk (deferred: a graph node in the reference), theta, (x_k, S_k, pi))."""
    N, D = x.shape
    if r_nk is None:
        g = torch.Generator(device='cpu').manual_seed(int(seed))
        e = -torch.log(torch.rand(N, K, generator=g, dtype=torch.float64))
        r_nk = (e / e.sum(1, keepdim=True)).to(device=x.device, dtype=x.dtype).contiguous()
    alpha_k, beta_k, m_k, C_k, v_k, x_k, S_k, pi = core.mixture_fit(x, r_nk, None, _prior_standard(K, D, seed, x.dtype, x.device))
    return r_nk, lazy_log(r_nk), (alpha_k, beta_k, m_k, C_k, v_k), (x_k, S_k, pi)


def fit(x, K, seed, nb_iters, *, r_nk=None):
    """The reference's driver loop (gmm.py:377-379: `for i in range(nb_iters): sess.run(update)`) as ONE call: nb_iters sweeps
    of `inference` on the state r_nk.  In fp32 with D <= 8, K <= 32 the responsibilities stay on chip between sweeps (the
    e-step kernel accumulates the next M-step's statistics); only the final state is written.  Same return as inference."""
    N, D = x.shape
    if r_nk is None:
        g = torch.Generator(device='cpu').manual_seed(int(seed))
        e = -torch.log(torch.rand(N, K, generator=g, dtype=torch.float64))
        r_nk = (e / e.sum(1, keepdim=True)).to(device=x.device, dtype=x.dtype).contiguous()
    alpha_k, beta_k, m_k, C_k, v_k, x_k, S_k, pi = core.mixture_fit(x, r_nk, None, _prior_standard(K, D, seed, x.dtype, x.device),
                                                                    n_sweeps=int(nb_iters))
    return r_nk, lazy_log(r_nk), (alpha_k, beta_k, m_k, C_k, v_k), (x_k, S_k, pi)
