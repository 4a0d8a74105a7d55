"""ctypes binding of libvmp_svae.so (include/vmp_svae.h).  There is no CPU fallback: every op needs the
CUDA library and CUDA tensors and fails loudly otherwise."""
import ctypes
import os

import torch

from . import _build

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, 'libvmp_svae.so')
_lib = None

c_int, c_i64, c_u64, c_dbl, c_ptr = ctypes.c_int, ctypes.c_int64, ctypes.c_uint64, ctypes.c_double, ctypes.c_void_p

# name -> argtypes (T-suffixed names are declared for both _f32 and _f64)
_SIGS_T = {
    'vmp_phi_prepare': [c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr],
    'vmp_theta_prepare_gauss': [c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr],
    'vmp_theta_prepare_student': [c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr],
    'vmp_svae_local_step': [c_i64, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_ptr, c_ptr, c_u64, c_i64,
                            c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, ctypes.c_size_t, c_ptr],
    'vmp_svae_small_step': [c_i64, c_int, c_int, c_int, c_int, c_int] + [c_ptr] * 8 + [c_dbl, c_ptr, c_ptr, c_ptr, c_u64, c_i64] +
                           [c_ptr] * 6 + [c_ptr],
    'vmp_svae_local_step_bwd': [c_i64, c_int, c_int, c_int] + [c_ptr] * 7 + [c_int, c_ptr, c_u64, c_ptr, c_ptr, c_ptr,
                                                                               c_dbl, c_ptr] + [c_ptr] * 6 +
                               [c_ptr, ctypes.c_size_t, c_ptr],
    'vmp_fill_noise': [c_i64, c_int, c_int, c_int, c_u64, c_i64, c_ptr, c_ptr, c_ptr],
    'vmp_suffstats': [c_i64, c_int, c_int, c_ptr, c_ptr, c_int, c_ptr, c_ptr, c_ptr],
    'vmp_suffstats_update': [c_i64, c_int, c_int, c_ptr, c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_dbl, c_ptr, c_int] + [c_ptr] * 10 + [c_ptr],
    'vmp_ng_update': [c_int, c_int, c_ptr, c_dbl, c_ptr, c_int] + [c_ptr] * 15 + [c_ptr],
    'vmp_mixture_mstep': [c_int, c_int, c_int, c_ptr] + [c_ptr] * 12 + [c_ptr],
    'vmp_mixture_estep': [c_i64, c_int, c_int] + [c_ptr] * 12 + [c_ptr],
    'vmp_mixture_fit': [c_i64, c_int, c_int, c_int, c_int] + [c_ptr] * 17 + [c_ptr, ctypes.c_size_t, c_ptr],
    'vmp_mixture_prepare': [c_int, c_int, c_int, c_ptr, c_ptr] + [c_ptr] * 17 + [c_ptr],
    'vmp_spd_inverse': [c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr],
    'vmp_decoder_loglike': [c_i64, c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr],
    'vmp_decoder_loglike_bwd': [c_i64, c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_dbl, c_ptr, c_ptr,
                                c_ptr, c_ptr],
    'vmp_decoder_metrics': [c_i64, c_int, c_int, c_int, c_int] + [c_ptr] * 8 + [c_ptr],
    'vmp_gaussian_sample_nat': [c_i64, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr],
    'vmp_gaussian_logprob_nat': [c_i64, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr],
}
_SIGS = {
    'vmp_version': [],
    'vmp_phi_record_len': [c_int],
    'vmp_theta_record_len': [c_int],
    'vmp_stats_len': [c_int],
    'vmp_svae_local_step_workspace_bytes': [c_int, c_int],
    'vmp_svae_local_step_bwd_workspace_bytes': [c_int, c_int],
    'vmp_mixture_fit_workspace_bytes': [c_int, c_int],
    'vmp_mixture_record_len': [c_int],
    'vmp_svae_small_step_supported': [c_i64, c_int, c_int],
    'vmp_svae_small_step_packed': [c_ptr],
    'vmp_mixture_estep_fused_f32': [c_i64, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_ptr],
    'vmp_fma_probe': [c_int, c_int, c_int, c_ptr, c_ptr],
}
EXPORTS = sorted([n for n in _SIGS] + [n + s for n in _SIGS_T for s in ('_f32', '_f64')])


class SmallStepArgs(ctypes.Structure):
    """VmpSmallStepArgs of include/vmp_svae.h"""
    _fields_ = [('N', c_i64), ('K', ctypes.c_int32), ('D', ctypes.c_int32), ('S', ctypes.c_int32), ('den_mode', ctypes.c_int32),
                ('only_alpha', ctypes.c_int32), ('dtype', ctypes.c_int32),
                ('eta1', c_ptr), ('eta2_diag', c_ptr), ('eta1_phi2', c_ptr), ('L_raw', c_ptr), ('pi_raw', c_ptr),
                ('theta', c_ptr * 5), ('prior', c_ptr * 5), ('theta_out', c_ptr * 5),
                ('rho', c_dbl), ('rho_dev', c_ptr), ('noise', c_ptr), ('gumbel_u', c_ptr), ('seed', c_u64),
                ('point_offset', c_i64), ('log_r', c_ptr), ('x_sample', c_ptr), ('z', c_ptr), ('x_k_samples', c_ptr),
                ('stats', c_ptr), ('elbo_acc', c_ptr), ('stream', c_ptr)]


class VmpError(RuntimeError):
    pass


def lib_path():
    return _LIB_PATH


def load(build_if_missing=True):
    """Load (building first if the .so is absent or stale and nvcc is available).  _build.build() is incremental: with
    an up-to-date .so it only compares mtimes; on a box without nvcc the shipped .so is used as is."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing and (not os.path.exists(_LIB_PATH) or _build.nvcc_available()):
        _build.build()
    if not os.path.exists(_LIB_PATH):
        raise VmpError('libvmp_svae.so is missing: run `python -m vmp_for_svae_b200._build` (needs nvcc); '
                       'there is no CPU fallback')
    lib = ctypes.CDLL(_LIB_PATH)
    for name, args in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = args, (ctypes.c_size_t if name.endswith('_bytes') else c_int)
    for name, args in _SIGS_T.items():
        for suf in ('_f32', '_f64'):
            fn = getattr(lib, name + suf, None)
            if fn is None:
                continue
            fn.argtypes, fn.restype = args, c_int
    _lib = lib
    return lib


_SUFFIX = {torch.float32: '_f32', torch.float64: '_f64'}


def suffix(dtype):
    try:
        return _SUFFIX[dtype]
    except KeyError:
        raise AssertionError('dtype must be float32 or float64, got %s' % dtype)


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    assert isinstance(t, torch.Tensor), 'expected a torch.Tensor'
    if not t.is_cuda:
        raise VmpError('vmp_for_svae_b200 ops need CUDA tensors (got device %s); there is no CPU path' % t.device)
    assert t.is_contiguous(), 'tensor must be contiguous'
    return c_ptr(t.data_ptr())


def stream_ptr(device=None):
    return c_ptr(torch.cuda.current_stream(device).cuda_stream)


def call(name, dtype, *args, device=None):
    """Call the _f32/_f64 flavour of `name`; non-zero status raises.  `device`: the CUDA device of the tensors — the
    launch is issued with that device current (the stream pointer passed in `args` belongs to it)."""
    lib = load()
    fn = getattr(lib, name + suffix(dtype))
    if device is not None and device.type == 'cuda' and device.index is not None \
            and device.index != torch.cuda.current_device():
        with torch.cuda.device(device):
            rc = fn(*args)
    else:
        rc = fn(*args)
    if rc != 0:
        if rc < 0:
            why = {-1: 'invalid argument', -2: 'latent dimension outside 1..64', -3: 'bad mode'}.get(rc, 'error')
            raise ValueError('%s: %s (status %d)' % (name, why, rc))
        raise VmpError('%s: CUDA error %d' % (name, rc))


def record_lens(D):
    lib = load()
    return lib.vmp_phi_record_len(D), lib.vmp_theta_record_len(D), lib.vmp_stats_len(D)


def as_f(t, dtype=None, device=None):
    """contiguous tensor of the working dtype on the working device"""
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if device is not None and t.device != device:
        t = t.to(device)
    return t.contiguous()
