"""distributions/niw.py of the reference (K-sized parameter algebra)."""
from .. import core


def _outer(a, b):
    """niw.py:46-49."""
    return a.unsqueeze(-1) * b.unsqueeze(-2)


def expected_values(niw_standard_params):
    """niw.py:8-17 : E[mu] = m ; E[Sigma] = inv(v * sym(inv(C))) (two batched SPD inverses, as the reference)."""
    beta, m, C, v = niw_standard_params
    C_inv, _ = core.spd_inverse(C, want_logdet=False)
    C_inv_sym = (C_inv + C_inv.transpose(-1, -2)) / 2.
    exp_C, _ = core.spd_inverse(C_inv_sym * v.unsqueeze(1).unsqueeze(2), want_logdet=False)
    return m.clone(), exp_C


def standard_to_natural(beta, m, C, v):
    """niw.py:20-30."""
    K, D = m.shape
    assert tuple(beta.shape) == (K,)
    b = beta.unsqueeze(-1) * m
    A = C + _outer(b, m)
    v_hat = v + D + 2
    return A, b, beta, v_hat


def natural_to_standard(A, b, beta, v_hat):
    """niw.py:33-43."""
    m = b / beta.unsqueeze(-1)
    K, D = m.shape
    assert tuple(beta.shape) == (K,)
    C = A - _outer(b, m)
    v = v_hat - D - 2
    return beta, m, C, v
