"""Kernel-only timing of vmp_svae_local_step for engine variants (tuning aid; run on the GPU box).

    python tools/time_local_step.py [--points 131072] [--K 128] [--D 64] [--variants 0,1,2,...]
Prints one line per variant: ms per launch, points/s, fraction of the nominal FP32 roofline, max |d log r| vs variant 0.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vmp_for_svae_b200 import core, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--points', type=int, default=131072)
    ap.add_argument('--K', type=int, default=128)
    ap.add_argument('--D', type=int, default=64)
    ap.add_argument('--S', type=int, default=1)
    ap.add_argument('--variants', default='0,1,2,3,4,5,6,7')
    ap.add_argument('--reps', type=int, default=3)
    ap.add_argument('--wide', default='0')
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    dt = torch.float32
    prior, theta, phi_gmm = synthetic.make_globals(a.K, a.D, seed=0, dtype=dt, device=dev)
    eta1, eta2d = synthetic.make_encoder_outputs(a.points, a.D, synthetic.cluster_centres(phi_gmm), seed=1, dtype=dt,
                                                 device=dev)
    phi_rec, theta_rec = core.phi_prepare(*phi_gmm), core.theta_prepare_gauss(theta)
    ws = core.local_step_workspace(a.K, a.D, dev)
    flops = a.points * a.K * (a.D ** 3 / 3.0 + (6 + 3 * a.S) * a.D * a.D)
    peak = 148 * 128 * 2 * 1.965e9
    ref = None
    for v in a.variants.split(','):
        os.environ['VMP_FAST_VARIANT'] = v.replace('g', '').replace('2d', '0')
        os.environ['VMP_FORCE_GENERIC'] = '1' if v == 'g' else '0'
        os.environ['VMP_FAST_WIDE'] = '1' if v == 'w' else '0'
        os.environ['VMP_FAST_2D'] = '1' if v.startswith('2d') else '0'
        os.environ['VMP_FAST_PF'] = v[2:] if v.startswith('pf') else '4'
        out = core.local_step(eta1, eta2d, phi_rec, theta_rec, a.S, seed=5, workspace=ws)   # warm-up
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            out = core.local_step(eta1, eta2d, phi_rec, theta_rec, a.S, seed=5, workspace=ws, log_r=out['log_r'],
                                  x_sample=out['x_sample'], z=out['z'])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        if ref is None:
            ref = out['log_r'].clone()
        diff = float((out['log_r'] - ref).abs().max())
        print('variant %-3s  %9.3f ms  %10.0f points/s  roofline %.3f  max|dlogr| %.2e  elbo %s'
              % (v, ms, a.points / ms * 1e3, flops / (ms * 1e-3) / peak, diff,
                 [round(float(t), 3) for t in out['elbo_acc'][:3]]), flush=True)


if __name__ == '__main__':
    main()
