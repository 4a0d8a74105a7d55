"""Kernel-only timing of vmp_suffstats on (a) the responsibilities the local step produces for the bench's synthetic
shard and (b) dense random responsibilities (no exact zeros), plus the fraction of exactly-zero weights.

    python tools/time_suffstats.py [--points 262144] [--K 128] [--D 64]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vmp_for_svae_b200 import core, synthetic  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--points', type=int, default=262144)
    ap.add_argument('--K', type=int, default=128)
    ap.add_argument('--D', type=int, default=64)
    a = ap.parse_args()
    dev, dt = torch.device('cuda', 0), torch.float32
    prior, theta, phi_gmm = synthetic.make_globals(a.K, a.D, seed=0, dtype=dt, device=dev)
    eta1, eta2d = synthetic.make_encoder_outputs(a.points, a.D, synthetic.cluster_centres(phi_gmm), seed=100, dtype=dt,
                                                 device=dev, spread=1.0)
    out = core.local_step(eta1, eta2d, core.phi_prepare(*phi_gmm), core.theta_prepare_gauss(theta), 1, seed=3)
    log_r, x = out['log_r'], out['x_sample']
    zero_frac = float((torch.exp(log_r) == 0).double().mean())
    stats = torch.zeros(a.K, core._lib.record_lens(a.D)[2], dtype=torch.float64, device=dev)
    ms_a = timed(lambda: core.suffstats(x, log_r, r_is_log=True, stats=stats))
    dense = torch.softmax(torch.randn(a.points, a.K, device=dev), dim=1).contiguous()
    ms_b = timed(lambda: core.suffstats(x, dense, r_is_log=False, stats=stats))
    fma = a.points * a.K * (a.D + 1) * (a.D + 2) / 2
    for name, ms in (('bench responsibilities (%.1f%% exact zeros)' % (100 * zero_frac), ms_a), ('dense responsibilities', ms_b)):
        print('%-48s %8.3f ms  %8.2f M points/s  %.1f TFLOP/s dense-equivalent' % (name, ms, a.points / ms * 1e-3, 2 * fma / ms * 1e-9))


if __name__ == '__main__':
    main()
