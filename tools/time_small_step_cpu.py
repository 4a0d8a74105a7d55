"""CPU-side cost of one call of the one-launch step (launch-bound shapes): wall time per call with the GPU kept busy vs idle."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vmp_for_svae_b200 import synthetic
from vmp_for_svae_b200.step import SVAEStep
for (N, K, D, S) in ((100, 10, 2, 10), (274, 10, 6, 10)):
    prior, theta, phi_gmm = synthetic.make_globals(K, D, seed=0, device='cuda')
    eta1, eta2d = synthetic.make_encoder_outputs(N, D, synthetic.cluster_centres(phi_gmm), seed=1, device='cuda')
    st = SVAEStep(N, K, D, S, device='cuda', use_dist=False)
    for i in range(50):
        st.step((eta1, eta2d), phi_gmm, theta, prior, 0.1, seed=i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(2000):
        st.step((eta1, eta2d), phi_gmm, theta, prior, 0.1, seed=i)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    b = st._bound
    t3 = time.perf_counter()
    for i in range(2000):
        b(0.1, seed=i)
    t4 = time.perf_counter()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(2000):
        b(0.1, seed=i)
    e1.record(); torch.cuda.synchronize()
    print((N, K, D, S), 'SVAEStep.step: %.2f us CPU per call (issue), drain %.2f us; bound call alone %.2f us CPU; GPU time per step %.2f us'
          % ((t1 - t0) / 2000 * 1e6, (t2 - t1) * 1e6 / 2000, (t4 - t3) / 2000 * 1e6, e0.elapsed_time(e1) / 2000 * 1e3))
