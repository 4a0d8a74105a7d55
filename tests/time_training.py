"""Training-iteration throughput of the SVAE driver (vmp_for_svae_b200.experiments) on the GPU next to the same
iteration built from the CPU oracle (oracle/backward.forward + torch autograd + Adam + oracle M-step / CVI), all host
threads.  Prints one JSON line.

    python tests/time_training.py [--dataset pinwheel|auto-like] [--iters 500] [--cpu-iters 30]
(lives under tests/ because it executes the oracle, which only test infrastructure may do)
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def cpu_iteration_rate(cfg, y_tr, M, S, iters):
    """One reference-style iteration on the CPU: encoder -> e_step + regulariser (oracle) -> decoder -> -elbo.backward()
    -> Adam; M-step + CVI update of theta with the oracle port.  fp32, all host threads."""
    from oracle import backward as ob, svae_port as sp
    from vmp_for_svae_b200 import experiments as ex
    torch.set_num_threads(os.cpu_count())
    K, L, U = cfg['K'], cfg['L'], cfg['U']
    dt = torch.float32
    enc = ex.ResNet(y_tr.shape[1], [(U, 'tanh'), (U, 'tanh'), (L, 'natparam')], 0.01, 0)
    dec = ex.ResNet(L, [(U, 'tanh'), (U, 'tanh'), (y_tr.shape[1], 'standard')], 0.01, 0)
    rs = np.random.RandomState(0)
    prior, theta = sp.init_mm(K, L, uniform=torch.as_tensor(rs.rand(K, L)), dtype=dt)
    phi = [torch.nn.Parameter(t.clone()) for t in sp.init_recognition_params(theta, K, normal=torch.as_tensor(rs.randn(K)).to(dt))]
    opt = torch.optim.Adam(list(enc.parameters()) + list(dec.parameters()) + phi, lr=cfg['lr'])
    g = torch.Generator().manual_seed(0)
    t0 = None
    for i in range(iters + 2):
        if i == 2:
            t0 = time.perf_counter()
        y = y_tr[torch.randint(0, y_tr.shape[0], (M,), generator=g)]
        e1, e2 = enc(y)
        W, m, cden = ob.theta_consts_gauss(theta)
        noise = torch.randn(M, K, L, S, generator=g, dtype=dt)
        x, log_r, reg = ob.forward(e1, e2, phi[0], phi[1], phi[2], W, m, cden, noise)
        means, vars_ = dec(x)
        elbo = sp.expected_diagonal_gaussian_loglike(y, means, vars_, weights=torch.exp(log_r)) - reg
        opt.zero_grad()
        (-elbo).backward()
        opt.step()
        with torch.no_grad():
            xs, _ = sp.subsample_x(x.detach(), log_r.detach(), u=torch.rand(M, S, generator=g, dtype=dt))
            star = sp.m_step(prior, xs[:, 0, :], torch.exp(log_r.detach()))
            sp.update_gmm_params(theta, star, cfg['lrcvi'])
    return iters / (time.perf_counter() - t0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--dataset', default='pinwheel')
    ap.add_argument('--iters', type=int, default=500)
    ap.add_argument('--cpu-iters', type=int, default=30)
    a = ap.parse_args()
    from vmp_for_svae_b200 import experiments as ex
    pin = a.dataset != 'auto-like'
    cfg = dict(dataset=a.dataset, method='svae-cvi', lr=0.01 if pin else 0.0003, lrcvi=0.1 if pin else 0.2,
               decay_rate=1.0 if pin else 0.95, K=10, L=2 if pin else 6, U=50, seed=0)
    M, S = (100 if pin else 64), 10
    X_tr, _, _, _ = ex.make_dataset(a.dataset)
    dev = torch.device('cuda', 0)
    y_tr = torch.as_tensor(X_tr, dtype=torch.float32, device=dev)
    tr = ex.SVAETrainer(cfg, y_tr.shape[1], device=dev, nb_samples=S)
    g = torch.Generator().manual_seed(0)
    for i in range(a.iters + 20):
        if i == 20:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        idx = torch.randint(0, y_tr.shape[0], (M,), generator=g).to(dev)
        tr.train_step(y_tr[idx].contiguous())
    torch.cuda.synchronize()
    gpu_rate = a.iters / (time.perf_counter() - t0)
    gt = ex.GraphedSVAETrainer(cfg, y_tr, M, device=dev, nb_samples=S).capture()
    for i in range(a.iters + 20):
        if i == 20:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        gt.train_step()
    torch.cuda.synchronize()
    graph_rate = a.iters / (time.perf_counter() - t0)
    cpu_rate = cpu_iteration_rate(cfg, y_tr.cpu(), M, S, a.cpu_iters)
    print(json.dumps({'workload': 'SVAE training iteration (%s: K=%d L=%d minibatch=%d S=%d, U=%d MLPs)' % (a.dataset, cfg['K'], cfg['L'], M, S, cfg['U']),
                      'gpu_iterations_per_s': gpu_rate, 'gpu_ms_per_iteration': 1e3 / gpu_rate,
                      'gpu_graph_iterations_per_s': graph_rate, 'gpu_graph_ms_per_iteration': 1e3 / graph_rate,
                      'cpu_iterations_per_s': cpu_rate, 'cpu_ms_per_iteration': 1e3 / cpu_rate, 'cpu_cores': os.cpu_count(),
                      'cpu_kind': 'oracle port of the training graph (torch-CPU fp32 autograd, all host threads)'}))


if __name__ == '__main__':
    main()
