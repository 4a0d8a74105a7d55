"""Timing of one VB-EM sweep of the standalone mixtures (BASELINE configs[2], C3: N=1M, K=32, D=8).

    python tools/time_mixture_sweep.py [--N 1000000] [--K 32] [--D 8] [--smm 1]
Reports ms per sweep (m_step statistics + M-step + SPD inverse + e_step), per kernel group, points/s and the
fraction of the HBM roofline (bytes/point = 4(D + 4K) for the SMM, 4(D + 2K) for the GMM)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vmp_for_svae_b200 import core  # noqa: E402
from vmp_for_svae_b200.models import gmm, smm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--N', type=int, default=1000000)
    ap.add_argument('--K', type=int, default=32)
    ap.add_argument('--D', type=int, default=8)
    ap.add_argument('--smm', type=int, default=1)
    ap.add_argument('--reps', type=int, default=10)
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    g = torch.Generator(device=dev).manual_seed(0)
    centres = 3.0 * torch.randn(7, a.D, generator=g, device=dev)
    x = centres[torch.randint(0, 7, (a.N,), generator=g, device=dev)] + torch.randn(a.N, a.D, generator=g, device=dev)
    x = ((x - x.mean(0)) / x.std(0)).contiguous()
    e = -torch.log(torch.rand(a.N, a.K, generator=g, device=dev))
    r = (e / e.sum(1, keepdim=True)).contiguous()
    u = torch.ones_like(r)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def sweep():
        if a.smm:
            return smm.inference(x, a.K, 5.0, seed=0, r_nk=r, u_nk=u)
        return gmm.inference(x, a.K, seed=0, r_nk=r)

    for _ in range(3):
        sweep()
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(a.reps):
        sweep()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    # the two N-sized kernels alone
    alpha_k = torch.ones(a.K, device=dev); beta_k = torch.ones(a.K, device=dev); v_k = torch.full((a.K,), a.D + 2.0, device=dev)
    m_k = centres.new_zeros(a.K, a.D); P_k = torch.eye(a.D, device=dev).repeat(a.K, 1, 1).contiguous()
    kap = torch.full((a.K,), 5.0, device=dev) if a.smm else None
    t = {}
    for name, fn in (('suffstats', lambda: core.suffstats(x, r, u_nk=u if a.smm else None)),
                     ('estep', lambda: core.mixture_estep(x, alpha_k, beta_k, m_k, P_k, v_k, kappa_k=kap, r=r, u_out=u if a.smm else None))):
        fn(); torch.cuda.synchronize()
        s0, s1 = ev(), ev()
        s0.record()
        for _ in range(a.reps):
            fn()
        s1.record(); torch.cuda.synchronize()
        t[name] = s0.elapsed_time(s1) / a.reps
    # multi-sweep driver loop in one call: r / u stay on chip between sweeps
    nfit = 8

    def fit():
        if a.smm:
            return smm.fit(x, a.K, 5.0, 0, nfit, r_nk=r, u_nk=u)
        return gmm.fit(x, a.K, 0, nfit, r_nk=r)
    fit(); torch.cuda.synchronize()
    f0, f1 = ev(), ev()
    f0.record()
    for _ in range(max(1, a.reps // 4)):
        fit()
    f1.record(); torch.cuda.synchronize()
    ms_fit = f0.elapsed_time(f1) / max(1, a.reps // 4) / nfit
    bpp = 4.0 * (a.D + (4 if a.smm else 2) * a.K)
    hbm = 6514.2e9
    flops = a.N * a.K * (4.0 * a.D * a.D + 4.0 * a.D)
    print(json.dumps({'shape': [a.N, a.K, a.D], 'smm': bool(a.smm), 'ms_per_sweep': ms, 'points_per_s': a.N / ms * 1e3,
                      'general_suffstats_ms': t['suffstats'], 'general_estep_ms': t['estep'],
                      'hbm_roofline_ms': a.N * bpp / hbm * 1e3, 'hbm_frac': a.N * bpp / hbm * 1e3 / ms,
                      'fp32_roofline_ms': flops / 74.45e12 * 1e3, 'fp32_frac': flops / 74.45e12 * 1e3 / ms,
                      'fit_ms_per_sweep': ms_fit, 'fit_sweeps': nfit, 'fit_fp32_frac': flops / 74.45e12 * 1e3 / ms_fit}))


if __name__ == '__main__':
    main()
