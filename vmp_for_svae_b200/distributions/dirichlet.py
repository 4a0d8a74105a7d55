"""distributions/dirichlet.py of the reference (K-sized; not performance relevant).

`expected_log_pi` is evaluated on the device inside the theta / e-step prologue kernels whenever it feeds the hot
path (prepare.cu, mixtures.cu); the stand-alone function below serves the public surface."""
import torch


def expected_log_pi(dir_standard_param):
    """dirichlet.py:8-12 : psi(alpha) - psi(sum alpha)."""
    a = dir_standard_param
    return torch.special.digamma(a) - torch.special.digamma(a.sum(dim=-1, keepdim=True))


def standard_to_natural(alpha):
    """dirichlet.py:15-17."""
    return alpha - 1


def natural_to_standard(alpha_nat):
    """dirichlet.py:20-22."""
    return alpha_nat + 1
