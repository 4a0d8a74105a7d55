"""Tiny invocations of the round-2 kernels for compute-sanitizer (memcheck / racecheck / synccheck), run on the GPU box:

    compute-sanitizer --tool racecheck python tools/sanitize_small.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vmp_for_svae_b200 import core, synthetic  # noqa: E402


def main():
    dev, dt = torch.device('cuda', 0), torch.float32
    for (N, K, D) in ((48, 8, 64), (96, 8, 32), (96, 8, 16), (40, 4, 24)):
        prior, theta, phi_gmm = synthetic.make_globals(K, D, seed=0, dtype=dt, device=dev)
        eta1, eta2d = synthetic.make_encoder_outputs(N, D, synthetic.cluster_centres(phi_gmm), seed=1, dtype=dt, device=dev)
        out = core.local_step(eta1, eta2d, core.phi_prepare(*phi_gmm), core.theta_prepare_gauss(theta), 1, seed=5)
        torch.cuda.synchronize()
        print('local_step', N, K, D, float(out['elbo_acc'][0]))
    g = torch.Generator().manual_seed(0)
    for (N, K, D) in ((700, 8, 32), (300, 12, 16), (1000, 8, 8), (333, 32, 8)):
        x = torch.randn(N, D, generator=g).to(dev)
        r = torch.softmax(torch.randn(N, K, generator=g), 1).to(dev).contiguous()
        u = (0.5 + torch.rand(N, K, generator=g)).to(dev).contiguous()
        s = core.suffstats(x, r, u_nk=u)
        torch.cuda.synchronize()
        print('suffstats', N, K, D, float(s[0, 0]))


if __name__ == '__main__':
    main()
