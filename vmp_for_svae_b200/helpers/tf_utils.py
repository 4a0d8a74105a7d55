"""helpers/tf_utils.py of the reference: only `logdet` is on the hot path (tf_utils.py:25-49)."""
from .. import core


def logdet(A, name='logdet'):
    """log(det(A)) for SPD A[..., D, D] = 2 * sum(log(diag(chol(A)))) — batched Cholesky kernel."""
    return core.spd_inverse(A, want_inv=False, want_logdet=True)[1]
