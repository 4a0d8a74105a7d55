#!/usr/bin/env python
"""bench.py — points/sec of the local VMP + natural-gradient step (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W           # this repo's CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N ...           # the reference's CPU path (oracle port of the TF1.3 graph)

One "step" = svae.e_step + subsample_x + ELBO regulariser + m_step + update_gmm_params over one batch of synthetic
encoder outputs.  Workload (config.workload): BASELINE.json configs[4] "C5" — K=128, D=64, S=1 — with the
N=2^26 points sharded 8 ways: every GPU owns 2^23 points (weak scaling; at --gpus 8 the job is exactly C5).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (K, D, S, points per GPU, rho)
    'c5': (128, 64, 1, 1 << 23, 0.2),
    'c4': (64, 32, 1, 1 << 16, 0.2),
    'c2': (10, 6, 10, 274, 0.2),
    'c1': (10, 2, 10, 100, 0.1),
}
METRIC = 'points_per_sec_local_vmp_plus_ng_step'


def flops_per_point(K, D, S):
    """SURVEY §8d canonical count: K (D^3/3 + (6+3S) D^2)."""
    return K * (D ** 3 / 3.0 + (6 + 3 * S) * D * D)


def bytes_per_point(K, D):
    """read eta1, eta2_diag; write log r, x_sample."""
    return 4.0 * (3 * D + K)


class ClockSampler(object):
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q,
                                       '--format=csv,noheader,nounits', '-lms', '200'],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except subprocess.TimeoutExpired:
            self.p.kill()
            out = self.p.communicate()[0]
        sm, mx, reasons, power = [], [], set(), []
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max(power) if power else None, 'samples': len(sm), 'reasons': sorted(reasons)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return json.load(f), 'measured'
    except (OSError, ValueError):
        return {'hbm_gbs': 6650.0}, 'fallback'


# ------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_rate(K, D, S, rho, sample_points, steps, warmup, seed=0):
    """Times the oracle port of the reference's TF1.3 graph (oracle/svae_port.svae_step: materialised [N,K,D,D],
    batched LU solves and Choleskys) in fp32 on all host threads, on `sample_points` points of the workload."""
    import numpy as np
    from oracle import svae_port
    torch.set_num_threads(os.cpu_count() or 1)
    dt = torch.float32
    rs = np.random.RandomState(seed)
    prior, theta = svae_port.init_mm(K, D, uniform=torch.as_tensor(rs.rand(K, D)), dtype=dt)
    mu_k, L_k, pi_k = svae_port.init_recognition_params(theta, K, normal=torch.as_tensor(rs.randn(K)))
    mu_k = mu_k + 0.1 * torch.as_tensor(rs.randn(K, D), dtype=dt)
    N = sample_points
    p1 = np.logaddexp(0.0, rs.randn(N, D))
    mu1 = mu_k.numpy()[rs.randint(0, K, N)] / 8.0 + 0.5 * rs.randn(N, D)
    phi_enc = (torch.as_tensor(mu1 * p1, dtype=dt), torch.as_tensor(-0.5 * p1, dtype=dt))
    noise = torch.as_tensor(rs.randn(N, K, D, S), dtype=dt)
    u = torch.as_tensor(rs.rand(N, S), dtype=dt)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        svae_port.svae_step(phi_enc, (mu_k, L_k, pi_k), [t.clone() for t in theta], prior, noise, u, rho)
        dtm = time.perf_counter() - t0
        if it >= warmup:
            times.append(dtm)
    mean = sum(times) / len(times)
    return N / mean, mean, torch.get_num_threads()


def run_reference(args):
    K, D, S, n_per_gpu, rho = WORKLOADS[args.workload]
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    sample = args.cpu_sample or max(8, min(256, (1 << 22) // (K * D * D // 16 + 1)))
    rate, sec, cores = cpu_reference_rate(K, D, S, rho, sample, args.steps, max(1, args.warmup))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': 'points/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args.workload), 'K': K, 'D': D, 'S': S},
        'cpu_baseline': {'value': rate, 'unit': 'points/s', 'cores': cores, 'kind': 'port',
                         'sample': '%d points of the workload per step (restatement of the TF1.3 graph in torch-CPU '
                                   'fp32, all host threads; TensorFlow 1.3 itself cannot run here)' % sample},
        'e2e': {'value': rate, 'unit': 'points/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))
    return 0


def workload_name(w):
    K, D, S, n, _ = WORKLOADS[w]
    if w == 'c5':
        return ('C5 large synthetic local-VMP scaling (BASELINE configs[4]): K=128, D=64, S=1, N=2^26 points sharded '
                '8 ways = 2^23 points per GPU (weak scaling; --gpus 8 is exactly C5)')
    return '%s: K=%d D=%d S=%d, %d points per GPU' % (w.upper(), K, D, S, n)


# ------------------------------------------------------------------------------------------- CUDA arm
def run_cuda(args):
    from vmp_for_svae_b200 import core, dist as vdist, synthetic
    from vmp_for_svae_b200.step import SVAEStep, svae_step_host
    rank, world, local = vdist.init_from_env('nccl')
    if world != args.gpus and rank == 0:
        print('warning: --gpus %d but WORLD_SIZE=%d' % (args.gpus, world), file=sys.stderr)
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device: there is no CPU fallback for the product path')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    K, D, S, n_per_gpu, rho = WORKLOADS[args.workload]
    if args.points:
        n_per_gpu = args.points
    dt = torch.float32
    prior, theta, phi_gmm = synthetic.make_globals(K, D, seed=0, dtype=dt, device=dev)
    centres = synthetic.cluster_centres(phi_gmm)
    eta1, eta2d = synthetic.make_encoder_outputs(n_per_gpu, D, centres, seed=100 + rank, dtype=dt, device=dev, spread=1.0)
    st = SVAEStep(n_per_gpu, K, D, S, dtype=dt, device=dev)
    theta0 = [t.clone() for t in theta]

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize(dev)

    ev = lambda: torch.cuda.Event(enable_timing=True)
    # ---- device-resident timing
    for i in range(args.warmup):
        st.step((eta1, eta2d), phi_gmm, theta, prior, rho, seed=i)
    sync_all()
    sampler = ClockSampler(local) if rank == 0 else None
    k_ev = [(ev(), ev()) for _ in range(args.steps)]
    t_start, t_end = ev(), ev()
    t_start.record()
    for i in range(args.steps):
        st.step((eta1, eta2d), phi_gmm, theta, prior, rho, seed=args.warmup + i, kernel_events=k_ev[i])
    t_end.record()
    sync_all()
    clocks = sampler.stop() if sampler else None
    ms_total = t_start.elapsed_time(t_end)
    ms_step = vdist.max_over_ranks(ms_total / args.steps, dev)
    ms_kernel = sum(a.elapsed_time(b) for a, b in k_ev) / args.steps
    bad = float(st.elbo_acc[3].item())

    # ---- end-to-end: pinned host inputs -> H2D -> step -> D2H of the step's results
    host = (torch.empty(eta1.shape, dtype=dt, pin_memory=True), torch.empty(eta2d.shape, dtype=dt, pin_memory=True))
    host[0].copy_(eta1); host[1].copy_(eta2d)
    for t, t0 in zip(theta, theta0):
        t.copy_(t0)
    # shards above 2^20 points go through the chunked pipeline: H2D of chunk c+1 on a second stream under the compute
    # of chunk c (two staging slots); small shards use one copy + one step
    chunk = (1 << 20) if n_per_gpu > (1 << 20) else None
    staging = None if chunk else (torch.empty_like(eta1), torch.empty_like(eta2d))
    for i in range(min(2, args.warmup)):
        svae_step_host(host, phi_gmm, theta, prior, rho, st, seed=i, staging=staging, chunk=chunk)
    sync_all()
    e0, e1 = ev(), ev()
    e0.record()
    for i in range(args.steps):
        elbo_h, alpha_h = svae_step_host(host, phi_gmm, theta, prior, rho, st, seed=args.warmup + i, staging=staging,
                                         chunk=chunk)
    e1.record()
    sync_all()
    ms_e2e = vdist.max_over_ranks(e0.elapsed_time(e1) / args.steps, dev)
    h2d = 2 * n_per_gpu * D * 4
    d2h = 4 * 8 + K * 4

    # ---- launch-bound shapes (C1 / C2): the same step replayed from a CUDA graph
    graph_replay = None
    if world == 1 and n_per_gpu * K * D * S <= (1 << 24):
        for t, t0 in zip(theta, theta0):
            t.copy_(t0)
        replay = st.make_graph((eta1, eta2d), phi_gmm, theta, prior, rho)
        for _ in range(args.warmup):
            replay()
        sync_all()
        g0, g1 = ev(), ev()
        g0.record()
        for _ in range(args.steps):
            replay()
        g1.record()
        sync_all()
        ms_graph = g0.elapsed_time(g1) / args.steps
        graph_replay = {'ms_per_step': ms_graph, 'value': n_per_gpu / (ms_graph * 1e-3), 'unit': 'points/s',
                        'note': 'whole step (prologues, injected torch-RNG noise, local step, statistics, update) '
                                'captured once in a CUDA graph'}

    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if rank != 0:
        return 0
    total_points = n_per_gpu * world
    peaks, peak_src = measured_peaks()
    fl = flops_per_point(K, D, S) * n_per_gpu
    by = bytes_per_point(K, D) * n_per_gpu
    fp32_nominal = 148 * 128 * 2 * 1.965e9 / 1e12
    # measured FP32 FMA peak of this GPU, same job (packed FFMA2 probe, tools/fp32_peak.py); falls back to nominal
    fp32_peak, fp32_src = fp32_nominal, 'nominal FP32 FMA peak 148 SM x 128 lanes x 2 x 1.965 GHz'
    try:
        sys.path.insert(0, os.path.join(ROOT, 'tools'))
        import fp32_peak as _probe
        m = max(_probe.measure(1), _probe.measure(0))
        if m > 0:
            fp32_peak = m
            fp32_src = ('measured in this job with the FFMA2 probe vmp_fma_probe (MEASURED_PEAKS.json has no FP32 '
                        'figure); nominal 148 SM x 128 lanes x 2 x 1.965 GHz = %.2f TFLOP/s' % fp32_nominal)
    except Exception:
        pass
    ach_tf = fl / (ms_kernel * 1e-3) / 1e12
    ach_gbs = by / (ms_kernel * 1e-3) / 1e9
    fp32_bound = fl / (fp32_nominal * 1e12) >= by / (peaks['hbm_gbs'] * 1e9)
    roofline = {
        'kernel': 'local_step (vmp_svae_local_step: per-pair Cholesky/solves + selected-sample pass)',
        'bound': 'fp32' if fp32_bound else 'hbm',
        'achieved': ach_tf if fp32_bound else ach_gbs,
        'peak': fp32_peak if fp32_bound else peaks['hbm_gbs'],
        'unit': 'TFLOP/s' if fp32_bound else 'GB/s',
        'frac': (ach_tf / fp32_peak) if fp32_bound else (ach_gbs / peaks['hbm_gbs']),
        'frac_of_nominal': (ach_tf / fp32_nominal) if fp32_bound else None,
        'peak_source': fp32_src if fp32_bound else ('hbm_gbs of MEASURED_PEAKS.json (%s)' % peak_src),
        'algorithmic_flops_per_launch': fl, 'algorithmic_bytes_per_launch': by,
        'kernel_ms': ms_kernel, 'kernel_share_of_step': ms_kernel / (ms_total / args.steps),
        'hbm_gbs_achieved': ach_gbs, 'hbm_frac': ach_gbs / peaks['hbm_gbs'], 'traffic': None,
    }
    # measured DRAM traffic of this kernel (one ncu capture, profiles/): bytes per point x points of one launch
    try:
        with open(os.path.join(ROOT, 'profiles', 'r1_traffic_local_step_fast64.json')) as f:
            tr = json.load(f)
        if (tr['K'], tr['D'], tr['S']) == (K, D, S):
            roofline['traffic'] = tr['dram_bytes_per_point'] * n_per_gpu
            roofline['traffic_source'] = tr['source']
    except (OSError, ValueError, KeyError):
        pass
    cpu = None
    if world == 1 or True:
        sample = args.cpu_sample or max(8, min(256, (1 << 22) // (K * D * D // 16 + 1)))
        try:
            rate, sec, cores = cpu_reference_rate(K, D, S, rho, sample, steps=2, warmup=1)
            cpu = {'value': rate, 'unit': 'points/s', 'cores': cores, 'kind': 'port',
                   'sample': '%d points of the workload, 2 timed steps after 1 warm-up (oracle port of the TF1.3 graph, '
                             'torch-CPU fp32, all host threads)' % sample}
        except Exception as e:  # the baseline leg must never take the GPU number down
            cpu = {'value': None, 'unit': 'points/s', 'cores': os.cpu_count(), 'kind': 'port', 'sample': 'failed: %r' % (e,)}
    line = {
        'metric': METRIC, 'value': total_points / (ms_step * 1e-3), 'unit': 'points/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args.workload), 'K': K, 'D': D, 'S': S, 'points_per_gpu': n_per_gpu,
                   'total_points': total_points, 'rho': rho,
                   'l2': 'inputs+outputs per step (%.1f MB/GPU) exceed the 126 MB L2; no explicit flush'
                         % ((by + 0.0) / 1e6) if by > 2.0e8 else 'working set fits L2 (launch-bound config)'},
        'clocks': clocks,
        'e2e': {'value': total_points / (ms_e2e * 1e-3), 'unit': 'points/s', 'ms_per_step': ms_e2e,
                'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'pipeline': ('chunks of %d points: H2D of chunk c+1 on a copy stream under the compute of chunk c' % chunk)
                if chunk else 'one copy, one step'},
        'gpu_launches': 6 * args.steps,
        'roofline': roofline,
        'cpu_baseline': cpu,
        'non_pd_pivots': bad,
    }
    if graph_replay is not None:
        line['graph_replay'] = graph_replay
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='cuda', choices=['cuda', 'reference'])
    ap.add_argument('--workload', default='c5', choices=sorted(WORKLOADS))
    ap.add_argument('--points', type=int, default=0, help='points per GPU (default: the workload\'s)')
    ap.add_argument('--cpu-sample', type=int, default=0, help='points per step of the CPU reference arm')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'cuda' else args.warmup
    if args.impl == 'reference':
        return run_reference(args)
    return run_cuda(args)


if __name__ == '__main__':
    sys.exit(main())
