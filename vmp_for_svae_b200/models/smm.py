"""Drop-in surface of the reference's models/smm.py (Bayesian Student-t mixture, Archambeau & Verleysen 2007).

Same kernels as models.gmm with the scale variables u_nk: moments weighted by r*u (suffstats.cu), M-step with the
1e-20 eps and v_k = v_0 + N_k (no +1), E-step with the linear-in-distance responsibilities of smm.py:119-128.
"""
import torch

from .. import core
from ..lazy import lazy_log
from ..distributions import dirichlet, niw
from . import svae


def update_Nk(r_nk):
    """smm.py:25-27."""
    N, K = r_nk.shape
    st = core.suffstats(torch.zeros(N, 1, dtype=r_nk.dtype, device=r_nk.device), r_nk)
    return st[:, 0].to(r_nk.dtype)


def update_Wk(ru_nk):
    """smm.py:30-32."""
    return update_Nk(ru_nk)


def update_xk(x, ru_nk, W_k, eps=1e-20):
    """smm.py:35-40."""
    st = core.suffstats(x, ru_nk)
    D = x.shape[1]
    return st[:, 2:2 + D].to(x.dtype) / (W_k.unsqueeze(1) + eps)


def update_Sk(x, ru_nk, W_k, x_k, eps=1e-20):
    """smm.py:43-50."""
    D = x.shape[1]
    st = core.suffstats(x, ru_nk)
    K = st.shape[0]
    s0, s1, s2 = st[:, 1].to(x.dtype), st[:, 2:2 + D].to(x.dtype), st[:, 2 + D:].reshape(K, D, D).to(x.dtype)
    S = s2 - x_k.unsqueeze(2) * s1.unsqueeze(1) - s1.unsqueeze(2) * x_k.unsqueeze(1) \
        + s0.reshape(-1, 1, 1) * x_k.unsqueeze(2) * x_k.unsqueeze(1)
    return S / (W_k.reshape(-1, 1, 1) + eps)


def update_alphak(alpha_0, N_k):
    """smm.py:53-55."""
    return alpha_0 + N_k


def update_betak(beta_0, W_k):
    """smm.py:58-60."""
    return beta_0 + W_k


def update_mk(beta_0, m_0, W_k, x_k, beta_k):
    """smm.py:63-71."""
    if beta_0.dim() == 1:
        beta_0 = beta_0.reshape(-1, 1)
    return (beta_0 * m_0 + W_k.unsqueeze(1) * x_k) / beta_k.unsqueeze(1)


def update_vk(v_0, N_k):
    """smm.py:74-76."""
    return v_0 + N_k


def update_Ck(C_0, x_k, W_k, m_0, beta_0, beta_k, S_k):
    """smm.py:79-85."""
    err = x_k - m_0
    return C_0 + W_k.reshape(-1, 1, 1) * S_k + (beta_0 * W_k / beta_k).reshape(-1, 1, 1) * err.unsqueeze(2) * err.unsqueeze(1)


def expct_log_pi(alpha_k):
    """smm.py:113-116."""
    return torch.special.digamma(alpha_k) - torch.special.digamma(alpha_k.sum())


def compute_expct_unk(expct_m_dist, kappa_k, D):
    """smm.py:131-137."""
    return (0.5 * (D + kappa_k)) / (0.5 * (expct_m_dist + kappa_k))


def e_step(x, alpha_k, beta_k, m_k, P_k, v_k, kappa_k, name='e_step'):
    """smm.py:140-164 -> (r_nk, u_nk, exp(E log pi))."""
    return core.mixture_estep(x, alpha_k, beta_k, m_k, P_k, v_k, kappa_k=kappa_k)


def m_step(x, r_nk, u_nk, alpha_0, beta_0, m_0, C_0, v_0, name='m_step'):
    """smm.py:167-196 -> (alpha_k, beta_k, m_k, C_k, v_k, x_k, S_k)."""
    N, D = x.shape
    stats = core.suffstats(x, r_nk, r_is_log=False, u_nk=u_nk)
    return core.mixture_mstep(stats, D, True, alpha_0, beta_0, m_0, C_0, v_0)


_PRIOR_CACHE = {}


def _prior_standard(K, D, seed, dtype, device):
    """The sweep's Dirichlet+NIW prior in standard parameters (smm.py: init_mm_params(...) -> niw.natural_to_standard,
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



def inference(x, K, kappa_init, seed, name='inference', *, r_nk=None, u_nk=None):
    """smm.py:199-245 : one VB-EM sweep; state (r_nk, u_nk) passed in / created, updated IN PLACE.
    Returns ((r_nk, u_nk), log_r_nk, theta, (x_k, S_k, pi))."""
    N, D = x.shape
    if r_nk is None:
        g = torch.Generator(device='cpu').manual_seed(int(seed))
        e = -torch.log(torch.rand(N, K, generator=g, dtype=torch.float64))
        r_nk = (e / e.sum(1, keepdim=True)).to(device=x.device, dtype=x.dtype).contiguous()
    if u_nk is None:
        u_nk = torch.ones(N, K, dtype=x.dtype, device=x.device)
    kappa_k = kappa_init * torch.ones(K, dtype=x.dtype, device=x.device)
    alpha_k, beta_k, m_k, C_k, v_k, x_k, S_k, pi = core.mixture_fit(x, r_nk, u_nk, _prior_standard(K, D, seed, x.dtype, x.device),
                                                                    kappa_k=kappa_k)
    return (r_nk, u_nk), lazy_log(r_nk), (alpha_k, beta_k, m_k, C_k, v_k, kappa_k), (x_k, S_k, pi)


def fit(x, K, kappa_init, seed, nb_iters, *, r_nk=None, u_nk=None):
    """The reference's driver loop (smm.py __main__: `for i in range(nb_iters): sess.run(update)`) as ONE call: nb_iters sweeps
    of `inference` on the state (r_nk, u_nk).  In fp32 with D <= 8, K <= 32 r and u stay on chip between sweeps (the e-step
    kernel accumulates the next M-step's statistics); only the final state is written.  Same return as inference."""
    N, D = x.shape
    if r_nk is None:
        g = torch.Generator(device='cpu').manual_seed(int(seed))
        e = -torch.log(torch.rand(N, K, generator=g, dtype=torch.float64))
        r_nk = (e / e.sum(1, keepdim=True)).to(device=x.device, dtype=x.dtype).contiguous()
    if u_nk is None:
        u_nk = torch.ones(N, K, dtype=x.dtype, device=x.device)
    kappa_k = kappa_init * torch.ones(K, dtype=x.dtype, device=x.device)
    alpha_k, beta_k, m_k, C_k, v_k, x_k, S_k, pi = core.mixture_fit(x, r_nk, u_nk, _prior_standard(K, D, seed, x.dtype, x.device),
                                                                    kappa_k=kappa_k, n_sweeps=int(nb_iters))
    return (r_nk, u_nk), lazy_log(r_nk), (alpha_k, beta_k, m_k, C_k, v_k, kappa_k), (x_k, S_k, pi)
