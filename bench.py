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
    # name: (K, D, S, points per GPU (weak scaling), rho, total points of the named config (strong scaling))
    'c5': (128, 64, 1, 1 << 23, 0.2, 1 << 26),
    'c4': (64, 32, 1, 1 << 16, 0.2, 1 << 16),
    'c3': (32, 8, 1, 1000000, 0.0, 1000000),        # SMM VB-EM sweep (smm.inference), kappa = 5
    'c2': (10, 6, 10, 274, 0.2, 274),
    'c1': (10, 2, 10, 100, 0.1, 100),
}
METRIC = 'points_per_sec_local_vmp_plus_ng_step'


def flops_per_point(K, D, S):
    """SURVEY §8d canonical count: K (D^3/3 + (6+3S) D^2)."""
    return K * (D ** 3 / 3.0 + (6 + 3 * S) * D * D)


def bytes_per_point(K, D):
    """read eta1, eta2_diag; write log r, x_sample."""
    return 4.0 * (3 * D + K)


def sweep_flops_per_point(K, D):
    """SURVEY §8d, GMM/SMM sweep: K (4 D^2 + 4 D)."""
    return K * (4.0 * D * D + 4.0 * D)


def sweep_bytes_per_point(K, D, smm=True):
    """read x, r[, u]; write r[, u]."""
    return 4.0 * (D + (4 if smm else 2) * K)


def count_kernel_launches(fn):
    """Kernel launches of one call of `fn`, counted by CUPTI through torch.profiler (outside the timed region).
    Returns (count, sorted kernel names) or (None, reason)."""
    try:
        from torch.profiler import ProfilerActivity, profile
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        names = [e.name for e in prof.events() if getattr(e, 'device_type', None) is not None and
                 str(e.device_type).endswith('CUDA') and 'memcpy' not in e.name.lower() and 'memset' not in e.name.lower()]
        if not names:
            return None, 'profiler returned no device events'
        short = sorted(set(n.split('(')[0].split('<')[0] for n in names))
        return len(names), short
    except Exception as e:          # CUPTI may be unavailable on a box: say so instead of guessing
        return None, 'torch.profiler unavailable: %r' % (e,)


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


def cpu_reference_sweep_rate(K, D, sample_points, steps, warmup, kappa=5.0, seed=0):
    """C3: the oracle port of smm.inference (oracle/mixtures.smm_sweep: the reference's einsum / [N,K,D,D] formulation)
    in fp32 on all host threads, on `sample_points` points."""
    import numpy as np
    from oracle import mixtures, svae_port
    torch.set_num_threads(os.cpu_count() or 1)
    dt = torch.float32
    rs = np.random.RandomState(seed)
    N = sample_points
    x = torch.as_tensor(rs.randn(7, D)[rs.randint(0, 7, N)] * 2.0 + rs.randn(N, D), dtype=dt)
    r = torch.as_tensor(rs.dirichlet(np.ones(K), N), dtype=dt)
    u = torch.ones(N, K, dtype=dt)
    prior = svae_port.init_mm_params(K, D, alpha_scale=0.05 / K, beta_scale=0.5, m_scale=0.0, C_scale=D + 0.5, v_init=D + 0.5,
                                     dtype=dt)
    kap = torch.full((K,), kappa, dtype=dt)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        r, u, _, _ = mixtures.smm_sweep(x, r, u, prior, kap)
        dtm = time.perf_counter() - t0
        if it >= warmup:
            times.append(dtm)
    mean = sum(times) / len(times)
    return N / mean, mean, torch.get_num_threads()


def cpu_sample_size(args, K, D):
    if args.workload == 'c3':
        return args.cpu_sample or 65536
    return args.cpu_sample or max(8, min(256, (1 << 22) // (K * D * D // 16 + 1)))


def cpu_rate(args, K, D, S, rho, sample, steps, warmup):
    if args.workload == 'c3':
        return cpu_reference_sweep_rate(K, D, sample, steps, warmup)
    return cpu_reference_rate(K, D, S, rho, sample, steps, warmup)


def run_reference(args):
    K, D, S, n_per_gpu, rho, n_total = WORKLOADS[args.workload]
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    sample = cpu_sample_size(args, K, D)
    rate, sec, cores = cpu_rate(args, K, D, S, rho, sample, args.steps, max(1, args.warmup))
    what = ('oracle port of smm.inference (einsum formulation of smm.py)' if args.workload == 'c3' else
            'restatement of the TF1.3 graph')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': 'points/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
        'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args.workload, args.scaling), 'K': K, 'D': D, 'S': S},
        'cpu_baseline': {'value': rate, 'unit': 'points/s', 'cores': cores, 'kind': 'port',
                         'sample': '%d points of the workload per step (%s in torch-CPU fp32, all host threads; TensorFlow '
                                   '1.3 itself cannot run here)' % (sample, what)},
        'e2e': {'value': rate, 'unit': 'points/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))
    return 0


def workload_name(w, scaling='weak'):
    K, D, S, n, _, n_total = WORKLOADS[w]
    if w == 'c5':
        if scaling == 'strong':
            return ('C5 large synthetic local-VMP scaling (BASELINE configs[4]): K=128, D=64, S=1, ALL N=2^26 points divided '
                    'over the ranks (strong scaling; --gpus 1 holds the whole config on one GPU)')
        return ('C5 large synthetic local-VMP scaling (BASELINE configs[4]): K=128, D=64, S=1, N=2^26 points sharded '
                '8 ways = 2^23 points per GPU (weak scaling; --gpus 8 is exactly C5)')
    if w == 'c3':
        return ('C3 SMM VB-EM sweep (BASELINE configs[2]): smm.inference, N=%d synthetic points %s, K=32, D=8, kappa=5; '
                'state (r,u) in / out' % (n_total if scaling == 'strong' else n, 'in total' if scaling == 'strong' else 'per GPU'))
    names = {'c4': 'C4 MNIST-shape GMM-SVAE (BASELINE configs[3])', 'c2': 'C2 Auto GMM-SVAE full batch (configs[1])',
             'c1': 'C1 pinwheel GMM-SVAE minibatch (configs[0])'}
    return '%s: K=%d D=%d S=%d, %d points %s' % (names[w], K, D, S, n_total if scaling == 'strong' else n,
                                                 'in total' if scaling == 'strong' else 'per GPU')


def replicas_identical(tensors, world, dev):
    """After the timed loop every rank must hold bit-identical global parameters (they all applied the same all-reduced
    statistics): compare against rank 0's copy, AND the verdicts over ranks."""
    if world == 1:
        return None
    same = 1
    for t in tensors:
        ref = t.clone()
        torch.distributed.broadcast(ref, src=0)
        bits = torch.int64 if t.element_size() == 8 else torch.int32        # bit patterns: NaNs must compare equal too
        same &= int(torch.equal(ref.view(bits), t.view(bits)))
    flag = torch.tensor([same], dtype=torch.int32, device=dev)
    torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
    return bool(flag.item())


def fp32_peak_measured():
    fp32_nominal = 148 * 128 * 2 * 1.965e9 / 1e12
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
    return fp32_peak, fp32_src, fp32_nominal


def cpu_baseline_leg(args, K, D, S, rho):
    if args.no_cpu_baseline:
        return {'value': None, 'unit': 'points/s', 'cores': 0, 'kind': 'port', 'sample': 'skipped (--no-cpu-baseline)'}
    sample = cpu_sample_size(args, K, D)
    try:
        rate, sec, cores = cpu_rate(args, K, D, S, rho, sample, steps=2, warmup=1)
        return {'value': rate, 'unit': 'points/s', 'cores': cores, 'kind': 'port',
                'sample': '%d points of the workload, 2 timed steps after 1 warm-up (oracle port of the reference graph, '
                          'torch-CPU fp32, all host threads)' % sample}
    except Exception as e:  # the baseline leg must never take the GPU number down
        return {'value': None, 'unit': 'points/s', 'cores': os.cpu_count(), 'kind': 'port', 'sample': 'failed: %r' % (e,)}


# ------------------------------------------------------------------------------------------- CUDA arm: SVAE step
def run_cuda(args):
    from vmp_for_svae_b200 import dist as vdist, synthetic
    from vmp_for_svae_b200.step import SVAEStep, svae_step_host
    rank, world, local = vdist.init_from_env('nccl')
    if world != args.gpus and rank == 0:
        print('warning: --gpus %d but WORLD_SIZE=%d' % (args.gpus, world), file=sys.stderr)
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device: there is no CPU fallback for the product path')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    K, D, S, n_per_gpu, rho, n_total = WORKLOADS[args.workload]
    if args.points:
        n_per_gpu = args.points
    offset = 0
    if args.scaling == 'strong':
        offset, end = vdist.shard_range(args.points or n_total, rank, world)
        n_per_gpu = end - offset
    else:
        offset = rank * n_per_gpu
    dt = torch.float32
    prior, theta, phi_gmm = synthetic.make_globals(K, D, seed=0, dtype=dt, device=dev)
    centres = synthetic.cluster_centres(phi_gmm)
    eta1, eta2d = synthetic.make_encoder_outputs(n_per_gpu, D, centres, seed=100 + rank, dtype=dt, device=dev, spread=1.0)
    st = SVAEStep(n_per_gpu, K, D, S, dtype=dt, device=dev, point_offset=offset)
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
        st.step((eta1, eta2d), phi_gmm, theta, prior, rho, seed=args.warmup + i,
                kernel_events=None if st.single_launch else k_ev[i])
    t_end.record()
    sync_all()
    clocks = sampler.stop() if sampler else None
    ms_total = t_start.elapsed_time(t_end)
    ms_step = vdist.max_over_ranks(ms_total / args.steps, dev)
    # one-launch step (C1 / C2): the kernel IS the step
    ms_kernel = ms_total / args.steps if st.single_launch else sum(a.elapsed_time(b) for a, b in k_ev) / args.steps
    bad = float(st.elbo_acc[3].item())
    identical = replicas_identical(theta, world, dev)
    launches, kernels = count_kernel_launches(lambda: st.step((eta1, eta2d), phi_gmm, theta, prior, rho, seed=12345))

    # ---- end-to-end: pinned host inputs -> H2D -> step -> D2H of the step's results
    e2e = None
    pinned_bytes = 2 * n_per_gpu * D * 4 + (n_per_gpu * (K + D) * 4 if args.e2e_outputs == 'all' else 0)
    if pinned_bytes > (24 << 30):
        e2e = {'value': None, 'unit': 'points/s', 'skipped': 'the shard needs %.1f GB of pinned host memory' % (pinned_bytes / 2 ** 30)}
    else:
        host = (torch.empty(eta1.shape, dtype=dt, pin_memory=True), torch.empty(eta2d.shape, dtype=dt, pin_memory=True))
        host[0].copy_(eta1); host[1].copy_(eta2d)
        back = None
        if args.e2e_outputs == 'all':
            back = (torch.empty(n_per_gpu, K, dtype=dt, pin_memory=True), torch.empty(n_per_gpu, D, dtype=dt, pin_memory=True))
        for t, t0 in zip(theta, theta0):
            t.copy_(t0)
        # shards above 2^20 points go through the chunked pipeline: H2D of chunk c+1 on a second stream under the compute of
        # chunk c (two staging slots); small shards use one copy + one step
        chunk = (1 << 20) if n_per_gpu > (1 << 20) else None
        staging = None if chunk else (torch.empty_like(eta1), torch.empty_like(eta2d))
        for i in range(min(2, args.warmup)):
            svae_step_host(host, phi_gmm, theta, prior, rho, st, seed=i, staging=staging, chunk=chunk, copy_back=back)
        sync_all()
        e0, e1 = ev(), ev()
        e0.record()
        for i in range(args.steps):
            res = svae_step_host(host, phi_gmm, theta, prior, rho, st, seed=args.warmup + i, staging=staging, chunk=chunk,
                                 copy_back=back)
        e1.record()
        sync_all()
        ms_e2e = vdist.max_over_ranks(e0.elapsed_time(e1) / args.steps, dev)
        h2d = 2 * n_per_gpu * D * 4
        theta_bytes = sum(t.numel() * t.element_size() for t in theta)
        d2h = 4 * 8 + theta_bytes + (n_per_gpu * (K + D) * 4 if back is not None else 0)
        e2e = {'value': n_per_gpu * world / (ms_e2e * 1e-3), 'unit': 'points/s', 'ms_per_step': ms_e2e,
               'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
               'd2h_contents': 'ELBO terms + the five updated global natural parameters (alpha, A, b, beta, v_hat)' +
                               (' + log_r[N,K] + x_sample[N,D]' if back is not None else
                                '; log_r / x_sample stay on the device (--e2e-outputs all copies them too)'),
               'pipeline': ('chunks of %d points: H2D of chunk c+1 on a copy stream under the compute of chunk c' % chunk)
               if chunk else 'one copy, one step'}

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
    total_points = n_per_gpu * world if args.scaling == 'weak' else (args.points or n_total)
    peaks, peak_src = measured_peaks()
    fl = flops_per_point(K, D, S) * n_per_gpu
    by = bytes_per_point(K, D) * n_per_gpu
    fp32_peak, fp32_src, fp32_nominal = fp32_peak_measured()
    ach_tf = fl / (ms_kernel * 1e-3) / 1e12
    ach_gbs = by / (ms_kernel * 1e-3) / 1e9
    fp32_bound = fl / (fp32_nominal * 1e12) >= by / (peaks['hbm_gbs'] * 1e9)
    roofline = {
        'kernel': ('svae_small_step_kernel (the whole step in one cluster launch)' if st.single_launch else
                   'local_step (vmp_svae_local_step: per-pair Cholesky/solves + selected-sample pass)'),
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
    # DRAM traffic of this kernel: STATIC figure from one ncu capture (profiles/), scaled to this launch's points — ncu cannot
    # run inside the timed job
    for name in ('r2_traffic_local_step_fast64.json', 'r1_traffic_local_step_fast64.json'):
        try:
            with open(os.path.join(ROOT, 'profiles', name)) as f:
                tr = json.load(f)
            if (tr['K'], tr['D'], tr['S']) == (K, D, S):
                roofline['traffic'] = tr['dram_bytes_per_point'] * n_per_gpu
                roofline['traffic_source'] = 'static: ' + tr['source']
                break
        except (OSError, ValueError, KeyError):
            pass
    line = {
        'metric': METRIC, 'value': total_points / (ms_step * 1e-3), 'unit': 'points/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True,
        'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args.workload, args.scaling), 'K': K, 'D': D, 'S': S, 'points_per_gpu': n_per_gpu,
                   'total_points': total_points, 'rho': rho,
                   'l2': 'inputs+outputs per step (%.1f MB/GPU) exceed the 126 MB L2; no explicit flush'
                         % ((by + 0.0) / 1e6) if by > 2.0e8 else 'working set fits L2 (launch-bound config)'},
        'clocks': clocks,
        'e2e': e2e,
        'gpu_launches': (launches * args.steps) if launches is not None else None,
        'gpu_launches_per_step': launches,
        'gpu_kernels': kernels,
        'roofline': roofline,
        'cpu_baseline': cpu_baseline_leg(args, K, D, S, rho),
        'non_pd_pivots': bad,
    }
    if identical is not None:
        line['replicas_bit_identical'] = identical
    if graph_replay is not None:
        line['graph_replay'] = graph_replay
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------- CUDA arm: C3 mixture sweep
def run_cuda_sweep(args):
    """BASELINE configs[2]: one SMM VB-EM sweep (smm.inference: M-step of the state (r,u) -> P = inv(C) -> e-step -> new
    state), points sharded across ranks, one all-reduce of the packed statistics per sweep."""
    from vmp_for_svae_b200 import dist as vdist
    from vmp_for_svae_b200.mixture_step import MixtureSweep
    from vmp_for_svae_b200.models import smm
    rank, world, local = vdist.init_from_env('nccl')
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device: there is no CPU fallback for the product path')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    K, D, _, n_per_gpu, _, n_total = WORKLOADS['c3']
    if args.points:
        n_per_gpu = args.points
    if args.scaling == 'strong':
        a, b = vdist.shard_range(args.points or n_total, rank, world)
        n_per_gpu = b - a
    dt = torch.float32
    kappa = 5.0
    g = torch.Generator(device=dev).manual_seed(17)                         # same mixture on every rank,
    centres = 3.0 * torch.randn(7, D, generator=g, device=dev)
    scales = 0.4 + 0.6 * torch.rand(7, D, generator=g, device=dev)
    g = torch.Generator(device=dev).manual_seed(100 + rank)                 # different points
    comp = torch.randint(0, 7, (n_per_gpu,), generator=g, device=dev)
    x = centres[comp] + scales[comp] * torch.randn(n_per_gpu, D, generator=g, device=dev)
    outl = torch.rand(n_per_gpu, generator=g, device=dev) < 0.05            # 5 % uniform outliers
    x[outl] = 12.0 * (torch.rand(int(outl.sum()), D, generator=g, device=dev) - 0.5)
    x = (x / 3.3).contiguous()
    e = -torch.log(torch.rand(n_per_gpu, K, generator=g, device=dev))
    r0 = (e / e.sum(1, keepdim=True)).contiguous()
    r, u = r0.clone(), torch.ones_like(r0)
    prior = smm._prior_standard(K, D, 0, dt, dev)
    kap = torch.full((K,), kappa, dtype=dt, device=dev)
    sw = MixtureSweep(K, D, prior, kappa_k=kap, dtype=dt, device=dev)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize(dev)

    ev = lambda: torch.cuda.Event(enable_timing=True)
    for _ in range(args.warmup):
        sw.sweep(x, r, u)
    sync_all()
    sampler = ClockSampler(local) if rank == 0 else None
    t0, t1 = ev(), ev()
    reps = max(args.steps, 20)                                               # a sweep is ~0.2 ms: time at least 20 of them
    t0.record()
    for _ in range(reps):
        out = sw.sweep(x, r, u)
    t1.record()
    sync_all()
    clocks = sampler.stop() if sampler else None
    ms_step = vdist.max_over_ranks(t0.elapsed_time(t1) / reps, dev)
    identical = replicas_identical([out['alpha_k'], out['m_k'], out['C_k']], world, dev)
    state_finite = bool(torch.isfinite(out['C_k']).all() and torch.isfinite(r).all() and torch.isfinite(u).all())
    launches, kernels = count_kernel_launches(lambda: sw.sweep(x, r, u))
    # multi-sweep driver loop in one call (r, u stay on chip between sweeps)
    nfit = 8
    sw.fit(x, r, u, nfit)
    sync_all()
    f0, f1 = ev(), ev()
    f0.record()
    for _ in range(max(2, reps // nfit)):
        sw.fit(x, r, u, nfit)
    f1.record()
    sync_all()
    ms_fit = vdist.max_over_ranks(f0.elapsed_time(f1) / max(2, reps // nfit) / nfit, dev)
    # ---- end to end: x from pinned host memory each sweep, the sweep's K-sized results read back
    host_x = torch.empty(x.shape, dtype=dt, pin_memory=True)
    host_x.copy_(x)
    xd = torch.empty_like(x)
    r.copy_(r0); u.fill_(1.0)
    keys = ('alpha_k', 'beta_k', 'm_k', 'C_k', 'v_k', 'x_k', 'S_k', 'pi')

    def e2e_step():
        xd.copy_(host_x, non_blocking=True)
        o = sw.sweep(xd, r, u)
        return [o[k].to('cpu') for k in keys]
    for _ in range(2):
        e2e_step()
    sync_all()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(reps):
        res = e2e_step()
    e1.record()
    sync_all()
    ms_e2e = vdist.max_over_ranks(e0.elapsed_time(e1) / reps, dev)
    d2h = sum(t.numel() * t.element_size() for t in res)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if rank != 0:
        return 0
    total_points = n_per_gpu * world if args.scaling == 'weak' else (args.points or n_total)
    peaks, peak_src = measured_peaks()
    by = sweep_bytes_per_point(K, D, True) * n_per_gpu
    fl = sweep_flops_per_point(K, D) * n_per_gpu
    fp32_nominal = 148 * 128 * 2 * 1.965e9 / 1e12
    ach_gbs = by / (ms_step * 1e-3) / 1e9
    roofline = {
        'kernel': 'one sweep = sweep_stats_mma_kernel (tensor cores) + sweep_prepare_kernel + sweep_estep_kernel (the two N-sized kernels are '
                  '>95 % of it; CUDA events around the whole sweep)',
        'bound': 'hbm', 'achieved': ach_gbs, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': ach_gbs / peaks['hbm_gbs'],
        'peak_source': 'hbm_gbs of MEASURED_PEAKS.json (%s)' % peak_src,
        'algorithmic_bytes_per_launch': by, 'algorithmic_flops_per_launch': fl, 'kernel_ms': ms_step,
        'kernel_share_of_step': 1.0,
        'fp32_tflops_achieved': fl / (ms_step * 1e-3) / 1e12, 'fp32_frac_of_nominal': fl / (ms_step * 1e-3) / 1e12 / fp32_nominal,
        'note': 'SURVEY 8d: the FP32 time of this config (124 us) exceeds its HBM time (84 us): both fractions are given',
        'traffic': None,
    }
    line = {
        'metric': METRIC, 'value': total_points / (ms_step * 1e-3), 'unit': 'points/s', 'n_gpus': world,
        'steps': reps, 'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True,
        'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name('c3', args.scaling), 'K': K, 'D': D, 'kappa': kappa, 'points_per_gpu': n_per_gpu,
                   'total_points': total_points,
                   'l2': 'x + r + u per sweep = %.0f MB per GPU > 126 MB L2; no explicit flush' % (by / 1e6)},
        'clocks': clocks,
        'e2e': {'value': total_points / (ms_e2e * 1e-3), 'unit': 'points/s', 'ms_per_step': ms_e2e,
                'h2d_bytes_per_step': n_per_gpu * D * 4, 'd2h_bytes_per_step': d2h,
                'd2h_contents': 'the sweep\'s K-sized results (alpha_k, beta_k, m_k, C_k, v_k, x_k, S_k, pi); the state r, u '
                                '[N,K] stays on the device between sweeps as the reference keeps it in tf.Variables'},
        'gpu_launches': (launches * reps) if launches is not None else None, 'gpu_launches_per_step': launches,
        'gpu_kernels': kernels,
        'roofline': roofline,
        'fit': {'ms_per_sweep': ms_fit, 'sweeps_per_call': nfit, 'value': total_points / (ms_fit * 1e-3), 'unit': 'points/s',
                'note': 'smm.fit: the reference\'s driver loop in one call; r, u are not written between sweeps'},
        'cpu_baseline': cpu_baseline_leg(args, K, D, 1, 0.0),
        'state_finite_after_timed_sweeps': state_finite,
    }
    if identical is not None:
        line['replicas_bit_identical'] = identical
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='cuda', choices=['cuda', 'reference'])
    ap.add_argument('--workload', default='c5', choices=sorted(WORKLOADS))
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='weak: the workload\'s points PER GPU; strong: the named config\'s total points divided over the ranks')
    ap.add_argument('--points', type=int, default=0, help='points per GPU (weak) / in total (strong); default: the workload\'s')
    ap.add_argument('--cpu-sample', type=int, default=0, help='points per step of the CPU reference arm')
    ap.add_argument('--no-cpu-baseline', action='store_true', help='skip the CPU leg (per-shape sweeps at several GPU counts)')
    ap.add_argument('--e2e-outputs', default='theta', choices=['theta', 'all'],
                    help='what the end-to-end leg copies back: ELBO + updated theta, or also log_r and x_sample')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'cuda' else args.warmup
    if args.impl == 'reference':
        return run_reference(args)
    if args.workload == 'c3':
        return run_cuda_sweep(args)
    return run_cuda(args)


if __name__ == '__main__':
    sys.exit(main())
