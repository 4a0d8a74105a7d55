"""Thin tensor-level wrappers over the C-ABI (one function per entry point of include/vmp_svae.h).

All tensors are CUDA, contiguous, float32 or float64 (the dtype of `eta1`/`x` picks the _f32/_f64 flavour).
Nothing here computes on the host or through torch ops: allocation of outputs only."""
import torch

from . import _lib
from ._lib import ptr, stream_ptr

DEN_GAUSS, DEN_STUDENT = 0, 1


def _chk(t, shape, dtype, name):
    assert isinstance(t, torch.Tensor), '%s must be a tensor' % name
    assert tuple(t.shape) == tuple(shape), '%s must have shape %s, got %s' % (name, tuple(shape), tuple(t.shape))
    assert t.dtype == dtype, '%s must be %s, got %s' % (name, dtype, t.dtype)
    return t.contiguous()


def phi_prepare(eta1_phi2, L_raw, pi_raw, out=None):
    """svae.unpack_recognition_gmm (svae.py:342-358) -> per-component records [K, phi_record_len(D)]."""
    K, D = eta1_phi2.shape
    dt, dev = eta1_phi2.dtype, eta1_phi2.device
    eta1_phi2 = _chk(eta1_phi2, (K, D), dt, 'eta1_phi2')
    L_raw = _chk(L_raw, (K, D, D), dt, 'L_k_raw')
    pi_raw = _chk(pi_raw, (K,), dt, 'pi_k_raw')
    plen = _lib.record_lens(D)[0]
    rec = out if out is not None else torch.empty(K, plen, dtype=dt, device=dev)
    _lib.call('vmp_phi_prepare', dt, K, D, ptr(eta1_phi2), ptr(L_raw), ptr(pi_raw), ptr(rec), stream_ptr(dev), device=dev)
    return rec


def theta_prepare_gauss(theta, out=None):
    """theta = (alpha, A, b, beta, v_hat) natural parameters -> records [K, theta_record_len(D)] (svae.py:204-208)."""
    alpha, A, b, beta, v_hat = theta
    K, D = b.shape
    dt, dev = b.dtype, b.device
    alpha = _chk(alpha, (K,), dt, 'alpha'); A = _chk(A, (K, D, D), dt, 'A'); b = _chk(b, (K, D), dt, 'b')
    beta = _chk(beta, (K,), dt, 'beta'); v_hat = _chk(v_hat, (K,), dt, 'v_hat')
    tlen = _lib.record_lens(D)[1]
    rec = out if out is not None else torch.empty(K, tlen, dtype=dt, device=dev)
    _lib.call('vmp_theta_prepare_gauss', dt, K, D, ptr(alpha), ptr(A), ptr(b), ptr(beta), ptr(v_hat), ptr(rec),
              stream_ptr(dev), device=dev)
    return rec


def theta_prepare_student(theta, out=None):
    """theta = (alpha_nat, mu_k, L_k_raw, dof) (experiments.py:174) -> records (svae.py:269-272; student_t.py:7-39)."""
    alpha, mu, L_raw, dof = theta
    K, D = mu.shape
    dt, dev = mu.dtype, mu.device
    alpha = _chk(alpha, (K,), dt, 'alpha'); mu = _chk(mu, (K, D), dt, 'mu_k')
    L_raw = _chk(L_raw, (K, D, D), dt, 'L_k'); dof = _chk(dof, (K,), dt, 'DoF')
    tlen = _lib.record_lens(D)[1]
    rec = out if out is not None else torch.empty(K, tlen, dtype=dt, device=dev)
    _lib.call('vmp_theta_prepare_student', dt, K, D, ptr(alpha), ptr(mu), ptr(L_raw), ptr(dof), ptr(rec),
              stream_ptr(dev), device=dev)
    return rec


def local_step_workspace(K, D, device):
    """Scratch the fast path of the local step needs (staged per-component records)."""
    lib = _lib.load()
    nbytes = int(lib.vmp_svae_local_step_workspace_bytes(K, D))
    return torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=device)


def local_step(eta1, eta2_diag, phi_rec, theta_rec, S, den_mode=DEN_GAUSS, noise=None, u=None, seed=0, x_in=None,
               want_x_sample=True, want_z=True, materialize_x_k=False, log_r=None, x_sample=None, z=None,
               elbo_acc=None, workspace=None, point_offset=0, use_engine=True):
    """The fused local step.  Returns dict(log_r, x_sample, z, x_k_samples, elbo_acc[4] double).
    u: uniforms[N,K] of the Gumbel-max categorical draw (None -> in-kernel Philox keyed by `seed`).
    point_offset: global index of eta1[0] (shards / host chunks draw the noise of the whole batch).
    use_engine=False withholds the workspace, i.e. runs the thread-per-pair kernels (tests compare the two)."""
    N, D = eta1.shape
    K = phi_rec.shape[0]
    dt, dev = eta1.dtype, eta1.device
    eta1 = _chk(eta1, (N, D), dt, 'eta1'); eta2_diag = _chk(eta2_diag, (N, D), dt, 'eta2_diag')
    plen, tlen, _ = _lib.record_lens(D)
    phi_rec = _chk(phi_rec, (K, plen), dt, 'phi_rec'); theta_rec = _chk(theta_rec, (K, tlen), dt, 'theta_rec')
    if noise is not None:
        noise = _chk(noise, (N, K, D, S), dt, 'noise')
    if u is not None:
        u = _chk(u, (N, K), dt, 'u (gumbel uniforms)')
    if x_in is not None:
        x_in = _chk(x_in, (N, K, S, D), dt, 'x_k_samps')
    log_r = log_r if log_r is not None else torch.empty(N, K, dtype=dt, device=dev)
    if want_x_sample and x_sample is None:
        x_sample = torch.empty(N, D, dtype=dt, device=dev)
    if want_z and z is None:
        z = torch.empty(N, dtype=torch.int32, device=dev)
    x_k = torch.empty(N, K, S, D, dtype=dt, device=dev) if materialize_x_k else None
    if elbo_acc is None:
        elbo_acc = torch.zeros(4, dtype=torch.float64, device=dev)
    if not use_engine:
        workspace = None
    elif workspace is None and dt == torch.float32 and x_in is None:
        workspace = local_step_workspace(K, D, dev)
    wbytes = workspace.numel() * workspace.element_size() if workspace is not None else 0
    _lib.call('vmp_svae_local_step', dt, N, K, D, S, ptr(eta1), ptr(eta2_diag), ptr(phi_rec), ptr(theta_rec),
              int(den_mode), ptr(noise), ptr(u), int(seed) & 0xFFFFFFFFFFFFFFFF, int(point_offset), ptr(x_in), ptr(log_r), ptr(x_sample),
              ptr(z), ptr(x_k), ptr(elbo_acc), ptr(workspace), wbytes, stream_ptr(dev), device=dev)
    return dict(log_r=log_r, x_sample=x_sample, z=z, x_k_samples=x_k, elbo_acc=elbo_acc)


def small_step_supported(N, K, D):
    return bool(_lib.load().vmp_svae_small_step_supported(int(N), int(K), int(D)))


def _ptr_array(tensors, n):
    import ctypes
    arr = (ctypes.c_void_p * n)()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr() if t is not None else None
    return arr


def small_step(eta1, eta2_diag, phi_gmm, theta, prior, theta_out, rho, S, den_mode=DEN_GAUSS, only_alpha=False, noise=None,
               u=None, seed=0, point_offset=0, log_r=None, x_sample=None, z=None, stats=None, elbo_acc=None,
               materialize_x_k=False):
    """The whole step in ONE launch (vmp_svae_small_step; D <= 8, K <= 32, N*K <= 8192, single GPU): prologues, local step,
    selection, statistics and the natural-gradient update of `theta_out` in place.  stats / elbo_acc are overwritten.
    Returns dict(log_r, x_sample, z, x_k_samples, elbo_acc, stats)."""
    N, D = eta1.shape
    K = phi_gmm[0].shape[0]
    dt, dev = eta1.dtype, eta1.device
    eta1 = _chk(eta1, (N, D), dt, 'eta1'); eta2_diag = _chk(eta2_diag, (N, D), dt, 'eta2_diag')
    e1, Lr, pr = (_chk(phi_gmm[0], (K, D), dt, 'eta1_phi2'), _chk(phi_gmm[1], (K, D, D), dt, 'L_k_raw'),
                  _chk(phi_gmm[2], (K,), dt, 'pi_k_raw'))
    for t in list(theta) + list(prior) + list(theta_out):
        assert t.is_cuda and t.is_contiguous() and t.dtype == dt, 'theta / prior tensors must be contiguous CUDA tensors'
    if noise is not None:
        noise = _chk(noise, (N, K, D, S), dt, 'noise')
    if u is not None:
        u = _chk(u, (N, K), dt, 'u (gumbel uniforms)')
    slen = _lib.record_lens(D)[2]
    log_r = log_r if log_r is not None else torch.empty(N, K, dtype=dt, device=dev)
    x_sample = x_sample if x_sample is not None else torch.empty(N, D, dtype=dt, device=dev)
    z = z if z is not None else torch.empty(N, dtype=torch.int32, device=dev)
    stats = stats if stats is not None else torch.empty(K, slen, dtype=torch.float64, device=dev)
    elbo_acc = elbo_acc if elbo_acc is not None else torch.empty(4, dtype=torch.float64, device=dev)
    x_k = torch.empty(N, K, S, D, dtype=dt, device=dev) if materialize_x_k else None
    rho_dev = None
    if isinstance(rho, torch.Tensor):
        rho_dev = _chk(rho.reshape(1), (1,), torch.float64, 'rho')
        rho = 0.0
    th, pa, to = _ptr_array(theta, 5), _ptr_array(prior, 5), _ptr_array(theta_out, 5)
    _lib.call('vmp_svae_small_step', dt, N, K, D, int(S), int(den_mode), int(bool(only_alpha)), ptr(eta1), ptr(eta2_diag),
              ptr(e1), ptr(Lr), ptr(pr), th, pa, to, float(rho), ptr(rho_dev), ptr(noise), ptr(u),
              int(seed) & 0xFFFFFFFFFFFFFFFF, int(point_offset), ptr(log_r), ptr(x_sample), ptr(z), ptr(x_k), ptr(stats),
              ptr(elbo_acc), stream_ptr(dev), device=dev)
    return dict(log_r=log_r, x_sample=x_sample, z=z, x_k_samples=x_k, elbo_acc=elbo_acc, stats=stats)


class BoundSmallStep(object):
    """vmp_svae_small_step with every pointer argument converted ONCE: the per-call cost of the launch-bound configurations is
    the Python / ctypes marshalling, not the kernel.  Rebind when any tensor is replaced (SVAEStep does that by comparing
    data_ptr()s); theta is updated in place, so a training loop binds once."""

    def __init__(self, eta1, eta2_diag, phi_gmm, theta, prior, theta_out, S, den_mode, only_alpha, log_r, x_sample, z, stats,
                 elbo_acc, point_offset=0):
        import ctypes
        N, D = eta1.shape
        K = phi_gmm[0].shape[0]
        dt, dev = eta1.dtype, eta1.device
        _chk(eta1, (N, D), dt, 'eta1'); _chk(eta2_diag, (N, D), dt, 'eta2_diag')
        _chk(phi_gmm[0], (K, D), dt, 'eta1_phi2'); _chk(phi_gmm[1], (K, D, D), dt, 'L_k_raw'); _chk(phi_gmm[2], (K,), dt, 'pi_k_raw')
        tensors = [eta1, eta2_diag] + list(phi_gmm) + list(theta) + list(prior) + list(theta_out)
        for t in tensors:
            assert t.is_cuda and t.is_contiguous() and t.dtype == dt, 'contiguous CUDA tensors of one dtype expected'
        self.key = tuple(t.data_ptr() for t in tensors)
        self.keep = tensors                                   # the bound tensors must stay alive
        self.N, self.K, self.D, self.S, self.dt, self.dev = N, K, D, int(S), dt, dev
        self.fn = _lib.load().vmp_svae_small_step_packed
        a = _lib.SmallStepArgs()
        a.N, a.K, a.D, a.S, a.den_mode, a.only_alpha = N, K, D, int(S), int(den_mode), int(bool(only_alpha))
        a.dtype = 0 if dt == torch.float32 else 1
        a.eta1, a.eta2_diag = eta1.data_ptr(), eta2_diag.data_ptr()
        a.eta1_phi2, a.L_raw, a.pi_raw = phi_gmm[0].data_ptr(), phi_gmm[1].data_ptr(), phi_gmm[2].data_ptr()
        for i, t in enumerate(theta):
            a.theta[i] = t.data_ptr()
        for i, t in enumerate(prior):
            a.prior[i] = t.data_ptr()
        for i, t in enumerate(theta_out):
            a.theta_out[i] = t.data_ptr()
        a.point_offset = int(point_offset)
        a.log_r, a.x_sample, a.z = log_r.data_ptr(), x_sample.data_ptr(), z.data_ptr()
        a.stats, a.elbo_acc = stats.data_ptr(), elbo_acc.data_ptr()
        self.args, self.ref = a, ctypes.byref(a)
        self.out = dict(log_r=log_r, x_sample=x_sample, z=z, x_k_samples=None, elbo_acc=elbo_acc, stats=stats)

    def __call__(self, rho, seed=0, noise=None, u=None):
        a = self.args
        if isinstance(rho, torch.Tensor):
            a.rho_dev, a.rho = rho.data_ptr(), 0.0
        else:
            a.rho_dev, a.rho = None, rho
        a.noise = noise.data_ptr() if noise is not None else None
        a.gumbel_u = u.data_ptr() if u is not None else None
        a.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        a.stream = torch.cuda.current_stream(self.dev).cuda_stream
        if self.dev.index is None or self.dev.index == torch.cuda.current_device():
            rc = self.fn(self.ref)
        else:
            with torch.cuda.device(self.dev):
                rc = self.fn(self.ref)
        if rc != 0:
            raise (ValueError if rc < 0 else _lib.VmpError)('vmp_svae_small_step: status %d' % rc)
        return self.out


def local_step_backward(eta1, eta2_diag, eta1_phi2, L_raw, pi_raw, phi_rec, theta_rec, S, log_r, gx, glr, greg,
                        den_mode=DEN_GAUSS, noise=None, seed=0, want_theta_rec_bar=False):
    """Reverse pass of the fused local step (vmp_svae_local_step_bwd): gradients of
    sum(gx * x_k_samples) + sum(glr * log_r) + greg * regulariser w.r.t. (eta1, eta2_diag, eta1_phi2, L_raw, pi_raw).
    `noise`/`seed` and `log_r` are those of the forward call; gx[N,K,S,D], glr[N,K].  With want_theta_rec_bar the
    gradient w.r.t. the theta record (W | m | cden) is appended (the SMM variant trains mu_k, L_k by gradient)."""
    N, D = eta1.shape
    K = phi_rec.shape[0]
    dt, dev = eta1.dtype, eta1.device
    eta1 = _chk(eta1, (N, D), dt, 'eta1'); eta2_diag = _chk(eta2_diag, (N, D), dt, 'eta2_diag')
    eta1_phi2 = _chk(eta1_phi2, (K, D), dt, 'eta1_phi2'); L_raw = _chk(L_raw, (K, D, D), dt, 'L_raw')
    pi_raw = _chk(pi_raw, (K,), dt, 'pi_raw')
    plen, tlen, _ = _lib.record_lens(D)
    phi_rec = _chk(phi_rec, (K, plen), dt, 'phi_rec'); theta_rec = _chk(theta_rec, (K, tlen), dt, 'theta_rec')
    log_r = _chk(log_r, (N, K), dt, 'log_r'); gx = _chk(gx, (N, K, S, D), dt, 'gx'); glr = _chk(glr, (N, K), dt, 'glr')
    if noise is not None:
        noise = _chk(noise, (N, K, D, S), dt, 'noise')
    greg_dev = None
    if isinstance(greg, torch.Tensor):          # device scalar: no host synchronisation (CUDA-graph friendly)
        greg_dev = greg.detach().reshape(1).to(device=dev, dtype=dt).contiguous()
        greg = 0.0
    lib = _lib.load()
    nbytes = int(lib.vmp_svae_local_step_bwd_workspace_bytes(K, D))
    work = torch.empty((nbytes + 7) // 8, dtype=torch.float64, device=dev)
    out = [torch.empty_like(t) for t in (eta1, eta2_diag, eta1_phi2, L_raw, pi_raw)]
    th_bar = torch.empty_like(theta_rec) if want_theta_rec_bar else None
    _lib.call('vmp_svae_local_step_bwd', dt, N, K, D, S, ptr(eta1), ptr(eta2_diag), ptr(eta1_phi2), ptr(L_raw),
              ptr(pi_raw), ptr(phi_rec), ptr(theta_rec), int(den_mode), ptr(noise), int(seed) & 0xFFFFFFFFFFFFFFFF,
              ptr(log_r), ptr(gx), ptr(glr), float(greg), ptr(greg_dev), ptr(out[0]), ptr(out[1]), ptr(out[2]), ptr(out[3]),
              ptr(out[4]), ptr(th_bar), ptr(work), nbytes, stream_ptr(dev), device=dev)
    return tuple(out) + ((th_bar,) if want_theta_rec_bar else ())


def fill_noise(N, K, D, S, seed, dtype, device, want_noise=True, want_u=True, point_offset=0):
    device = torch.device(device)
    noise = torch.empty(N, K, D, S, dtype=dtype, device=device) if want_noise else None
    u = torch.empty(N, K, dtype=dtype, device=device) if want_u else None
    _lib.call('vmp_fill_noise', dtype, N, K, D, S, int(seed) & 0xFFFFFFFFFFFFFFFF, int(point_offset), ptr(noise), ptr(u),
              stream_ptr(device), device=device)
    return noise, u


def suffstats(x, r, r_is_log=False, u_nk=None, stats=None):
    """stats[K, D*D+D+2] (double) += [sum r, sum w, sum w x, sum w x x^T]."""
    N, D = x.shape
    K = r.shape[1]
    dt, dev = x.dtype, x.device
    x = _chk(x, (N, D), dt, 'x'); r = _chk(r, (N, K), dt, 'r_nk')
    if u_nk is not None:
        u_nk = _chk(u_nk, (N, K), dt, 'u_nk')
    slen = _lib.record_lens(D)[2]
    if stats is None:
        stats = torch.zeros(K, slen, dtype=torch.float64, device=dev)
    _lib.call('vmp_suffstats', dt, N, K, D, ptr(x), ptr(r), int(bool(r_is_log)), ptr(u_nk), ptr(stats), stream_ptr(dev), device=dev)
    return stats


def suffstats_update(x, r, stats, counter, rho, prior, theta, r_is_log=False, u_nk=None, only_alpha=False):
    """Statistics + natural-gradient update in one launch (vmp_suffstats_update; single-GPU path): stats += [...] and the last
    CTA applies theta <- (1-rho) theta + rho (prior + stats terms) in place.  counter: 1-element int32 CUDA tensor, zero."""
    N, D = x.shape
    K = r.shape[1]
    dt, dev = x.dtype, x.device
    x = _chk(x, (N, D), dt, 'x'); r = _chk(r, (N, K), dt, 'r_nk')
    rho_dev = None
    if isinstance(rho, torch.Tensor):
        rho_dev = _chk(rho.reshape(1), (1,), torch.float64, 'rho')
        rho = 0.0
    assert counter.dtype == torch.int32 and counter.numel() == 1 and counter.is_cuda
    if only_alpha:
        pp = [ptr(prior[0].contiguous()), None, None, None, None]
        tp = [ptr(theta[0]), None, None, None, None]
    else:
        for t in theta:
            assert t.is_contiguous() and t.dtype == dt
        pp = [ptr(t.contiguous()) for t in prior]
        tp = [ptr(t) for t in theta]
    _lib.call('vmp_suffstats_update', dt, N, K, D, ptr(x), ptr(r), int(bool(r_is_log)), ptr(u_nk), ptr(stats), ptr(counter),
              float(rho), ptr(rho_dev), int(bool(only_alpha)), *pp, *tp, stream_ptr(dev), device=dev)
    return stats


def ng_update(stats, rho, prior, theta, only_alpha=False, want_star=False):
    """theta <- (1-rho) theta + rho (prior + stats terms), in place.  Returns theta* when want_star.
    rho: python float, or a 1-element float64 CUDA tensor (device-resident step size, CUDA-graph friendly)."""
    rho_dev = None
    if isinstance(rho, torch.Tensor):
        rho_dev = _chk(rho.reshape(1), (1,), torch.float64, 'rho')
        rho = 0.0
    alpha = theta[0]
    K = alpha.shape[0]
    dt, dev = alpha.dtype, alpha.device
    if only_alpha:
        D = (stats.shape[1] and int(round((-1 + (1 + 4 * (stats.shape[1] - 2)) ** 0.5) / 2)))
        star = [torch.empty_like(alpha)] if want_star else [None]
        _lib.call('vmp_ng_update', dt, K, D, ptr(stats), float(rho), ptr(rho_dev), 1, ptr(prior[0].contiguous()), None, None, None,
                  None, ptr(alpha), None, None, None, None, ptr(star[0]), None, None, None, None, stream_ptr(dev), device=dev)
        return star if want_star else None
    D = theta[2].shape[1]
    for t in theta:
        assert t.is_contiguous() and t.dtype == dt
    p = [t.contiguous() for t in prior]
    star = [torch.empty_like(t) for t in theta] if want_star else [None] * 5
    _lib.call('vmp_ng_update', dt, K, D, ptr(stats), float(rho), ptr(rho_dev), 0, ptr(p[0]), ptr(p[1]), ptr(p[2]), ptr(p[3]), ptr(p[4]),
              ptr(theta[0]), ptr(theta[1]), ptr(theta[2]), ptr(theta[3]), ptr(theta[4]),
              ptr(star[0]), ptr(star[1]), ptr(star[2]), ptr(star[3]), ptr(star[4]), stream_ptr(dev), device=dev)
    return star if want_star else None


def mixture_mstep(stats, D, is_smm, alpha_0, beta_0, m_0, C_0, v_0):
    K = stats.shape[0]
    dt, dev = m_0.dtype, m_0.device
    alpha_0 = _chk(alpha_0, (K,), dt, 'alpha_0'); beta_0 = _chk(beta_0, (K,), dt, 'beta_0')
    m_0 = _chk(m_0, (K, D), dt, 'm_0'); C_0 = _chk(C_0, (K, D, D), dt, 'C_0'); v_0 = _chk(v_0, (K,), dt, 'v_0')
    e = lambda *s: torch.empty(*s, dtype=dt, device=dev)
    out = (e(K), e(K), e(K, D), e(K, D, D), e(K), e(K, D), e(K, D, D))
    _lib.call('vmp_mixture_mstep', dt, K, D, int(bool(is_smm)), ptr(stats), ptr(alpha_0), ptr(beta_0), ptr(m_0), ptr(C_0),
              ptr(v_0), *[ptr(o) for o in out], stream_ptr(dev), device=dev)
    return out


def mixture_estep(x, alpha_k, beta_k, m_k, P_k, v_k, kappa_k=None, missing_mask=None, r=None, u_out=None, work=None):
    N, D = x.shape
    K = alpha_k.shape[0]
    dt, dev = x.dtype, x.device
    x = _chk(x, (N, D), dt, 'x'); alpha_k = _chk(alpha_k, (K,), dt, 'alpha_k'); beta_k = _chk(beta_k, (K,), dt, 'beta_k')
    m_k = _chk(m_k, (K, D), dt, 'm_k'); P_k = _chk(P_k, (K, D, D), dt, 'P_k'); v_k = _chk(v_k, (K,), dt, 'v_k')
    if kappa_k is not None:
        kappa_k = _chk(kappa_k, (K,), dt, 'kappa_k')
    if missing_mask is not None:
        missing_mask = _chk(missing_mask.to(torch.uint8), (N, D), torch.uint8, 'missing_data_mask')
    r = r if r is not None else torch.empty(N, K, dtype=dt, device=dev)
    if kappa_k is not None and u_out is None:
        u_out = torch.empty(N, K, dtype=dt, device=dev)
    pi = torch.empty(K, dtype=dt, device=dev)
    work = work if work is not None else torch.empty(K, dtype=dt, device=dev)
    _lib.call('vmp_mixture_estep', dt, N, K, D, ptr(x), ptr(alpha_k), ptr(beta_k), ptr(m_k), ptr(P_k), ptr(v_k),
              ptr(kappa_k), ptr(missing_mask), ptr(r), ptr(u_out), ptr(pi), ptr(work), stream_ptr(dev), device=dev)
    return r, u_out, pi


def spd_inverse(mats, want_inv=True, want_logdet=True):
    """Batched SPD inverse / logdet over the leading dims of mats[..., D, D]."""
    D = mats.shape[-1]
    lead = mats.shape[:-2]
    dt, dev = mats.dtype, mats.device
    flat = mats.reshape(-1, D, D).contiguous()
    B = flat.shape[0]
    inv = torch.empty_like(flat) if want_inv else None
    ld = torch.empty(B, dtype=dt, device=dev) if want_logdet else None
    _lib.call('vmp_spd_inverse', dt, B, D, ptr(flat), ptr(inv), ptr(ld), stream_ptr(dev), device=dev)
    return (inv.reshape(*lead, D, D) if want_inv else None), (ld.reshape(lead) if want_logdet else None)


def decoder_loglike(y, means, out2, w, mode):
    """acc (double scalar tensor) of the weighted decoder reduction; mode 0 gaussian, 1 bernoulli."""
    N, K, S, Dobs = out2.shape
    dt, dev = out2.dtype, out2.device
    y = _chk(y, (N, Dobs), dt, 'y'); out2 = out2.contiguous(); w = _chk(w, (N, K), dt, 'weights')
    if means is not None:
        means = _chk(means, (N, K, S, Dobs), dt, 'means')
    acc = torch.zeros(1, dtype=torch.float64, device=dev)
    _lib.call('vmp_decoder_loglike', dt, N, K, S, Dobs, int(mode), ptr(y), ptr(means), ptr(out2), ptr(w), ptr(acc),
              stream_ptr(dev), device=dev)
    return acc


def decoder_loglike_backward(y, means, out2, w, mode, scale):
    """(g_means, g_out2, g_w) = scale * d acc / d (means, out2, w) of `decoder_loglike`."""
    N, K, S, Dobs = out2.shape
    dt, dev = out2.dtype, out2.device
    y = _chk(y, (N, Dobs), dt, 'y'); out2 = out2.contiguous(); w = _chk(w, (N, K), dt, 'weights')
    g_means = None
    if means is not None:
        means = _chk(means, (N, K, S, Dobs), dt, 'means')
        g_means = torch.empty_like(means)
    g_out2, g_w = torch.empty_like(out2), torch.empty_like(w)
    _lib.call('vmp_decoder_loglike_bwd', dt, N, K, S, Dobs, int(mode), ptr(y), ptr(means), ptr(out2), ptr(w),
              float(scale), ptr(g_means), ptr(g_out2), ptr(g_w), stream_ptr(dev), device=dev)
    return g_means, g_out2, g_w


def decoder_metrics(y, means, out2, mode, target=None, mask=None, log_w_nks=None, want_sq=True, want_lse=True):
    """(sq[N,K], lse[N,K]) of vmp_decoder_metrics (see include/vmp_svae.h); mode 0 gaussian, 1 bernoulli."""
    N, K, S, Dobs = out2.shape
    dt, dev = out2.dtype, out2.device
    y = _chk(y, (N, Dobs), dt, 'y'); out2 = out2.contiguous(); means = _chk(means, (N, K, S, Dobs), dt, 'means')
    if target is not None:
        target = _chk(target, (N, Dobs), dt, 'target')
    if mask is not None:
        mask = _chk(mask.to(torch.uint8), (N, Dobs), torch.uint8, 'missing_data_mask')
    if log_w_nks is not None:
        log_w_nks = _chk(log_w_nks, (N, K, S), dt, 'log_weights')
    sq = torch.empty(N, K, dtype=dt, device=dev) if want_sq else None
    lse = torch.empty(N, K, dtype=dt, device=dev) if want_lse else None
    _lib.call('vmp_decoder_metrics', dt, N, K, S, Dobs, int(mode), ptr(y), ptr(target), ptr(means), ptr(out2), ptr(mask),
              ptr(log_w_nks), ptr(sq), ptr(lse), stream_ptr(dev), device=dev)
    return sq, lse


def gaussian_logprob_nat(x, eta1, eta2, log_w=None, per_samp=False):
    N, K, D = eta1.shape
    dt, dev = eta1.dtype, eta1.device
    eta1 = eta1.contiguous(); eta2 = _chk(eta2, (N, K, D, D), dt, 'eta2')
    if per_samp:
        S = x.shape[2]
        x = _chk(x, (N, K, S, D), dt, 'x_samps')
        out = torch.empty(N, K, S, dtype=dt, device=dev)
    else:
        S = 0
        x = _chk(x, (N, D), dt, 'x')
        out = torch.empty(N, K, dtype=dt, device=dev)
    if log_w is not None:
        log_w = _chk(log_w, (K,), dt, 'log weights')
    _lib.call('vmp_gaussian_logprob_nat', dt, N, K, S, D, ptr(x), ptr(eta1), ptr(eta2), ptr(log_w), ptr(out),
              stream_ptr(dev), device=dev)
    return out


def gaussian_sample_nat(eta1, eta2, noise):
    """x[N,K,S,D] = P^-1 eta1 + L^-T noise for dense eta1[N,K,D(,1)], eta2[N,K,D,D], noise[N,K,D,S] (svae.py:95-119).
    Returns (x, non_pd) with non_pd a device int32 count of non-positive-definite systems."""
    N, K, D, _ = eta2.shape
    S = noise.shape[-1]
    dt, dev = eta2.dtype, eta2.device
    eta1 = _chk(eta1.reshape(N, K, D), (N, K, D), dt, 'eta1'); eta2 = _chk(eta2, (N, K, D, D), dt, 'eta2')
    noise = _chk(noise, (N, K, D, S), dt, 'noise')
    x = torch.empty(N, K, S, D, dtype=dt, device=dev)
    bad = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.call('vmp_gaussian_sample_nat', dt, N * K, D, S, ptr(eta1), ptr(eta2), ptr(noise), ptr(x), ptr(bad),
              stream_ptr(dev), device=dev)
    return x, bad


def mixture_fit(x, r, u, prior_std, kappa_k=None, n_sweeps=1):
    """n_sweeps VB-EM sweeps of gmm.inference / smm.inference on the state (r[, u]) IN PLACE (vmp_mixture_fit).
    prior_std = (alpha_0, beta_0, m_0, C_0, v_0) in standard parameters.  Returns the last M-step's
    (alpha_k, beta_k, m_k, C_k, v_k, x_k, S_k, pi)."""
    N, D = x.shape
    K = r.shape[1]
    dt, dev = x.dtype, x.device
    x = _chk(x, (N, D), dt, 'x')
    assert r.is_contiguous() and tuple(r.shape) == (N, K) and r.dtype == dt, 'r_nk must be a contiguous [N,K] state tensor'
    is_smm = kappa_k is not None
    if is_smm:
        assert u is not None and u.is_contiguous() and tuple(u.shape) == (N, K) and u.dtype == dt, 'u_nk state'
        kappa_k = _chk(kappa_k, (K,), dt, 'kappa_k')
    alpha_0, beta_0, m_0, C_0, v_0 = prior_std
    alpha_0 = _chk(alpha_0, (K,), dt, 'alpha_0'); beta_0 = _chk(beta_0, (K,), dt, 'beta_0')
    m_0 = _chk(m_0, (K, D), dt, 'm_0'); C_0 = _chk(C_0, (K, D, D), dt, 'C_0'); v_0 = _chk(v_0, (K,), dt, 'v_0')
    e = lambda *s: torch.empty(*s, dtype=dt, device=dev)
    out = (e(K), e(K), e(K, D), e(K, D, D), e(K), e(K, D), e(K, D, D), e(K))
    lib = _lib.load()
    nbytes = int(lib.vmp_mixture_fit_workspace_bytes(K, D))
    work = torch.empty((nbytes + 7) // 8, dtype=torch.float64, device=dev)
    _lib.call('vmp_mixture_fit', dt, N, K, D, int(is_smm), int(n_sweeps), ptr(x), ptr(alpha_0), ptr(beta_0), ptr(m_0),
              ptr(C_0), ptr(v_0), ptr(kappa_k), ptr(r), ptr(u) if is_smm else None, *[ptr(o) for o in out], ptr(work), nbytes,
              stream_ptr(dev), device=dev)
    return out


def mixture_record_len(D):
    return int(_lib.load().vmp_mixture_record_len(int(D)))


def mixture_prepare(stats, prior_std, is_smm, kappa_k=None, stats_next=None, out=None, want_general=False, rec=None):
    """K-sized phase of a sweep (vmp_mixture_prepare): M-step in standard parameters from `stats` + P_k = C_k^-1 + e-step
    constants.  Returns dict(alpha_k, beta_k, m_k, C_k, v_k, x_k, S_k, pi[, P_k, cst][, rec])."""
    alpha_0, beta_0, m_0, C_0, v_0 = prior_std
    K, D = m_0.shape
    dt, dev = m_0.dtype, m_0.device
    e = lambda *s: torch.empty(*s, dtype=dt, device=dev)
    o = out if out is not None else dict(alpha_k=e(K), beta_k=e(K), m_k=e(K, D), C_k=e(K, D, D), v_k=e(K), x_k=e(K, D),
                                         S_k=e(K, D, D), pi=e(K))
    if want_general and 'P_k' not in o:
        o['P_k'], o['cst'] = e(K, D, D), e(K)
    if rec is None and not want_general:
        rec = torch.empty(K, mixture_record_len(D), dtype=torch.float32, device=dev)
    if rec is not None:
        o['rec'] = rec
    _lib.call('vmp_mixture_prepare', dt, K, D, int(bool(is_smm)), ptr(stats), ptr(stats_next), ptr(alpha_0), ptr(beta_0), ptr(m_0),
              ptr(C_0), ptr(v_0), ptr(kappa_k), ptr(o['alpha_k']), ptr(o['beta_k']), ptr(o['m_k']), ptr(o['C_k']), ptr(o['v_k']),
              ptr(o['x_k']), ptr(o['S_k']), ptr(o['pi']), ptr(o.get('P_k')), ptr(o.get('cst')), ptr(o.get('rec')),
              stream_ptr(dev), device=dev)
    return o


def mixture_estep_fused(x, rec, is_smm, r=None, u=None, stats_next=None, write_state=True):
    """fp32, D <= 8, K <= 32 e-step from packed records; optionally accumulates the statistics of the new state."""
    N, D = x.shape
    K = rec.shape[0]
    dev = x.device
    assert x.dtype == torch.float32 and rec.dtype == torch.float32
    lib = _lib.load()
    rc = lib.vmp_mixture_estep_fused_f32(N, K, D, int(bool(is_smm)), ptr(x), ptr(rec), ptr(r), ptr(u), ptr(stats_next),
                                         int(bool(write_state)), stream_ptr(dev))
    if rc != 0:
        raise (ValueError if rc < 0 else _lib.VmpError)('vmp_mixture_estep_fused_f32: status %d' % rc)
