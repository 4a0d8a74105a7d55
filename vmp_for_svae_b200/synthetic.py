"""Synthetic inputs of the hot path: encoder potentials of the shape the reference's encoder emits
(eta2_diag = -1/2 softplus(raw), vae.py:40; eta1 = mu1 * (-2 eta2_diag) with mu1 drawn around K cluster centres)
and global parameters from the reference's own initialisers (svae.init_mm / init_recognition_params).
Plain tensor generation; used by tests, smoke() and bench.py."""
import torch

from .models import svae


def make_globals(K, D, seed=0, dtype=torch.float32, device='cuda', perturb=0.1):
    """(prior, theta, phi_gmm) as the reference initialises them, optionally nudged away from the symmetric init so
    that L_k has off-diagonal entries and pi_k is non-uniform (every code path is exercised)."""
    prior, theta = svae.init_mm(K, D, seed=seed, param_device=device, dtype=dtype)
    mu_k, L_k, pi_k = svae.init_recognition_params(theta, K, seed=seed)
    if perturb:
        g = torch.Generator(device='cpu').manual_seed(1234 + seed)
        mu_k = mu_k + perturb * torch.randn(K, D, generator=g, dtype=torch.float64).to(device=device, dtype=dtype)
        L_k = L_k + (perturb / D ** 0.5) * torch.randn(K, D, D, generator=g, dtype=torch.float64).to(device=device, dtype=dtype)
        pi_k = pi_k + perturb * torch.randn(K, generator=g, dtype=torch.float64).to(device=device, dtype=dtype)
    return prior, theta, (mu_k.contiguous(), L_k.contiguous(), pi_k.contiguous())


def cluster_centres(phi_gmm):
    """mu2_k = P2_k^-1 eta1_k of the recognition GMM (where its components sit in latent space)."""
    _, eta2, _ = svae.unpack_recognition_gmm(phi_gmm)
    P2 = -2.0 * eta2
    return torch.linalg.solve(P2.double(), phi_gmm[0].double().unsqueeze(-1)).squeeze(-1).to(phi_gmm[0].dtype)


def make_encoder_outputs(N, D, centres, seed=0, dtype=torch.float32, device='cuda', spread=1.0, chunk=1 << 20):
    """eta1[N,D], eta2_diag[N,D]; mu1 = centre[c_n] + spread * N(0, 1), c_n uniform over the K centres."""
    K = centres.shape[0]
    g = torch.Generator(device=device).manual_seed(int(seed))
    eta1 = torch.empty(N, D, dtype=dtype, device=device)
    eta2d = torch.empty(N, D, dtype=dtype, device=device)
    for s in range(0, N, chunk):
        e = min(N, s + chunk)
        raw = torch.randn(e - s, D, generator=g, dtype=dtype, device=device)
        p1 = torch.nn.functional.softplus(raw)
        idx = torch.randint(0, K, (e - s,), generator=g, device=device)
        mu1 = centres[idx] + spread * torch.randn(e - s, D, generator=g, dtype=dtype, device=device)
        eta2d[s:e] = -0.5 * p1
        eta1[s:e] = mu1 * p1
    return eta1, eta2d
