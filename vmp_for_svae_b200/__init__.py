"""vmp_for_svae_b200 — B200-native (sm_100a) implementation of the data-parallel hot path of
emtiyaz/vmp-for-svae: the per-point local VMP step (encoder Gaussian potentials x GMM/SMM global factors)
and the natural-gradient (CVI) global update it drives.

Layout mirrors the reference's module paths so that it is a drop-in for that path:
    vmp_for_svae_b200.models.svae / .gmm / .smm      <- models/svae.py, models/gmm.py, models/smm.py
    vmp_for_svae_b200.distributions.{gaussian,niw,dirichlet,student_t}
    vmp_for_svae_b200.helpers.tf_utils.logdet
plus `step.SVAEStep` (the fused step used by bench.py) and `core` (one wrapper per C-ABI entry point).
Everything computes in libvmp_svae.so (hand-written CUDA); there is no CPU or eager fallback.
"""
from . import _lib  # noqa: F401

__all__ = ['core', 'step', 'models', 'distributions', 'helpers']
