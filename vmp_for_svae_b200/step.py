"""The fused hot path: one `SVAEStep.step()` = local VMP step over this rank's shard of points + global
natural-gradient update (experiments.py:208-260 in one allocation-free sequence of kernel launches):

    phi_prepare, theta_prepare  (K-sized prologues)
    local_step                  log r, selected sample x[n, z_n, 0], ELBO-regulariser partials
    suffstats                   [N_k, sum r x, sum r x x^T] of this shard, double
    all-reduce (NCCL)           the packed K*(D^2+D+2)+4 doubles — the only exchange of the step
    ng_update                   theta <- (1-rho) theta + rho (prior + stats)   (identical on every rank)
On one GPU the last two lines disappear: the update runs in the tail of the statistics kernel (vmp_suffstats_update).

Points are sharded contiguously across ranks, phi_gmm / theta / prior are replicated.  The ELBO is evaluated with
theta BEFORE the update (the reference leaves the order undefined; the oracle fixes the same order).
"""
import torch

from . import core, _lib


class SVAEStep(object):
    def __init__(self, N_local, K, D, S=1, dtype=torch.float32, device='cuda', den_mode=core.DEN_GAUSS,
                 process_group=None, use_dist=None, point_offset=0):
        """point_offset: global index of this rank's first point (dist.shard_range(...)[0]).  The in-kernel noise is
        keyed by the global pair index, so the sharded step draws exactly what the single-GPU step over the whole
        batch draws (every rank passes the same `seed`)."""
        self.N, self.K, self.D, self.S = int(N_local), int(K), int(D), int(S)
        self.point_offset = int(point_offset)
        self.dtype, self.device, self.den_mode = dtype, torch.device(device), den_mode
        plen, tlen, slen = _lib.record_lens(D)
        e = lambda *s, dt=dtype: torch.empty(*s, dtype=dt, device=self.device)
        self.phi_rec, self.theta_rec = e(K, plen), e(K, tlen)
        self.log_r, self.x_sample = e(self.N, K), e(self.N, D)
        self.z = e(self.N, dt=torch.int32)
        # one contiguous double buffer: [stats K*slen | elbo 4] -> a single all-reduce
        self.red = torch.zeros(K * slen + 4, dtype=torch.float64, device=self.device)
        self.stats = self.red[:K * slen].view(K, slen)
        self.elbo_acc = self.red[K * slen:]
        self.workspace = core.local_step_workspace(K, D, self.device) if dtype == torch.float32 else None
        self.pg = process_group
        if use_dist is None:
            use_dist = torch.distributed.is_available() and torch.distributed.is_initialized() and \
                torch.distributed.get_world_size(process_group) > 1
        self.use_dist = bool(use_dist)
        self.counter = torch.zeros(1, dtype=torch.int32, device=self.device)   # ticket counter of the fused NG tail
        # launch-bound shapes on one GPU (C1 / C2): the whole step is ONE kernel on a thread-block cluster
        self.single_launch = (not self.use_dist) and core.small_step_supported(self.N, K, D)
        self._bound, self._bound_calls = None, 0

    def step(self, phi_enc, phi_gmm, theta, prior, rho, seed=0, noise=None, u=None, only_alpha=False,
             kernel_events=None):
        """Runs one step; mutates `theta` in place; returns dict(log_r, x_sample, z, elbo_acc) (device tensors,
        valid on the current stream).  elbo_acc = [sum r*num, sum r*den, regulariser, #bad pivots] over ALL ranks."""
        eta1, eta2_diag = phi_enc
        if self.single_launch and kernel_events is None:
            pr, to = ([prior[0]], [theta[0]]) if only_alpha else (prior, theta)
            b = self._bound
            if b is not None:
                # same tensor OBJECTS as last time?  (theta is updated in place: a loop binds once.)  The bound call keeps
                # references to its tensors, so an id() cannot be recycled while it is bound; storage swaps are caught by
                # re-checking the data pointers every 64th call.
                ts = (eta1, eta2_diag) + tuple(phi_gmm) + tuple(theta) + tuple(pr) + tuple(to)
                ids = tuple(map(id, ts))
                self._bound_calls += 1
                if ids != b.ids or b.only_alpha != bool(only_alpha) or \
                        ((self._bound_calls & 63) == 0 and tuple(t.data_ptr() for t in ts) != b.key):
                    b = None
            if b is None:
                b = core.BoundSmallStep(eta1, eta2_diag, phi_gmm, theta, pr, to, self.S, self.den_mode, only_alpha, self.log_r,
                                        self.x_sample, self.z, self.stats, self.elbo_acc, point_offset=self.point_offset)
                b.only_alpha = bool(only_alpha)
                b.ids = tuple(map(id, (eta1, eta2_diag) + tuple(phi_gmm) + tuple(theta) + tuple(pr) + tuple(to)))
                self._bound = b
            if noise is not None:
                noise = core._chk(noise, (self.N, self.K, self.D, self.S), self.dtype, 'noise')
            if u is not None:
                u = core._chk(u, (self.N, self.K), self.dtype, 'u (gumbel uniforms)')
            return b(rho, seed=seed, noise=noise, u=u)
        core.phi_prepare(phi_gmm[0], phi_gmm[1], phi_gmm[2], out=self.phi_rec)
        if self.den_mode == core.DEN_GAUSS:
            core.theta_prepare_gauss(theta, out=self.theta_rec)
        else:
            core.theta_prepare_student(theta, out=self.theta_rec)
        self.red.zero_()
        if kernel_events is not None:       # CUDA events bracketing the dominant kernel (bench.py roofline)
            kernel_events[0].record()
        core.local_step(eta1, eta2_diag, self.phi_rec, self.theta_rec, self.S, den_mode=self.den_mode, noise=noise,
                        u=u, seed=seed, log_r=self.log_r, x_sample=self.x_sample, z=self.z, elbo_acc=self.elbo_acc,
                        workspace=self.workspace, point_offset=self.point_offset)
        if kernel_events is not None:
            kernel_events[1].record()
        if not self.use_dist:
            # single GPU: the natural-gradient update rides in the tail of the statistics reduction (one launch)
            core.suffstats_update(self.x_sample, self.log_r, self.stats, self.counter, rho,
                                  [prior[0]] if only_alpha else prior, [theta[0]] if only_alpha else theta, r_is_log=True,
                                  only_alpha=only_alpha)
        else:
            core.suffstats(self.x_sample, self.log_r, r_is_log=True, stats=self.stats)
            torch.distributed.all_reduce(self.red, group=self.pg)
            if only_alpha:
                core.ng_update(self.stats, rho, [prior[0]], [theta[0]], only_alpha=True)
            else:
                core.ng_update(self.stats, rho, prior, theta)
        return dict(log_r=self.log_r, x_sample=self.x_sample, z=self.z, elbo_acc=self.elbo_acc)


    def make_graph(self, phi_enc, phi_gmm, theta, prior, rho):
        """Capture the whole step in a CUDA graph for the launch-bound shapes (C1 / C2: a few hundred points).  The
        tensors passed here are the graph's static inputs (update them in place between replays); noise and Gumbel
        uniforms are drawn inside the graph by torch's graph-safe generator and injected, rho lives in a device scalar
        (`self.rho_dev`, update in place for a decaying schedule).  Returns `replay() -> dict` (same dict as step())."""
        assert not self.use_dist, 'graph capture of the multi-rank step is not supported'
        assert self.N * self.K * self.D * self.S <= (1 << 24), 'injected-noise graph is meant for small shards'
        self.rho_dev = torch.full((1,), float(rho), dtype=torch.float64, device=self.device)

        def body():
            noise = torch.randn(self.N, self.K, self.D, self.S, dtype=self.dtype, device=self.device)
            u = torch.rand(self.N, self.K, dtype=self.dtype, device=self.device)
            return self.step(phi_enc, phi_gmm, theta, prior, self.rho_dev, noise=noise, u=u)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        saved = [t.clone() for t in theta]
        with torch.cuda.stream(side):
            for _ in range(2):
                body()
        torch.cuda.current_stream(self.device).wait_stream(side)
        for t, t0 in zip(theta, saved):                                  # the warm-up steps must not move theta
            t.copy_(t0)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = body()
        self._graph = graph

        def replay():
            graph.replay()
            return out
        return replay


def svae_step_host(phi_enc_host, phi_gmm, theta, prior, rho, stepper, seed=0, staging=None, chunk=None, noise=None,
                   u=None, copy_back=None):
    """End-to-end form of the step for HOST-resident encoder outputs (pinned CPU tensors): copies this rank's
    eta1 / eta2_diag to the device, runs the step, and reads the step's results back to the host: the ELBO terms and all
    five updated global natural parameters; with copy_back=(log_r_host[N,K], x_sample_host[N,D]) (pinned) also the per-point
    outputs.  Returns (elbo_terms ndarray[4], [alpha, A, b, beta, v_hat] as CPU tensors).

    chunk=None: one copy, then `stepper.step`.  chunk=C (points): the shard is processed in chunks of C points with the
    host->device copy of chunk c+1 (own stream, two staging slots) overlapping the local step + statistics of chunk c;
    the statistics and ELBO terms accumulate on the device and the all-reduce / natural-gradient update run once at
    the end, so the result is the same VMP step: the in-kernel noise stream is keyed by the global pair index
    (point_offset of each chunk), so chunked, unchunked and sharded calls make the same draws (tests)."""
    eta1_h, eta2_h = phi_enc_host
    N = eta1_h.shape[0]
    st = stepper
    if chunk is None or chunk >= N:
        if staging is None:
            staging = (torch.empty(eta1_h.shape, dtype=eta1_h.dtype, device=st.device),
                       torch.empty(eta2_h.shape, dtype=eta2_h.dtype, device=st.device))
        staging[0].copy_(eta1_h, non_blocking=True)
        staging[1].copy_(eta2_h, non_blocking=True)
        out = st.step(staging, phi_gmm, theta, prior, rho, seed=seed, noise=noise, u=u)
    else:
        chunk = int(chunk)
        if getattr(st, '_chunk_state', None) is None or st._chunk_state[0] != chunk:
            slots = [tuple(torch.empty(chunk, st.D, dtype=st.dtype, device=st.device) for _ in range(2)) for _ in range(2)]
            st._chunk_state = (chunk, slots, torch.cuda.Stream(device=st.device),
                               [torch.cuda.Event() for _ in range(2)], [torch.cuda.Event() for _ in range(2)])
        _, slots, cstream, ready, free = st._chunk_state
        cur = torch.cuda.current_stream(st.device)
        core.phi_prepare(phi_gmm[0], phi_gmm[1], phi_gmm[2], out=st.phi_rec)
        if st.den_mode == core.DEN_GAUSS:
            core.theta_prepare_gauss(theta, out=st.theta_rec)
        else:
            core.theta_prepare_student(theta, out=st.theta_rec)
        st.red.zero_()
        for ev in free:
            ev.record(cur)
        nchunks = (N + chunk - 1) // chunk
        for c in range(nchunks):
            lo, hi = c * chunk, min(N, (c + 1) * chunk)
            m, slot = hi - lo, c & 1
            with torch.cuda.stream(cstream):
                cstream.wait_event(free[slot])                       # the step of chunk c-2 is done with this slot
                slots[slot][0][:m].copy_(eta1_h[lo:hi], non_blocking=True)
                slots[slot][1][:m].copy_(eta2_h[lo:hi], non_blocking=True)
                ready[slot].record(cstream)
            cur.wait_event(ready[slot])
            core.local_step(slots[slot][0][:m], slots[slot][1][:m], st.phi_rec, st.theta_rec, st.S, den_mode=st.den_mode,
                            noise=None if noise is None else noise[lo:hi], u=None if u is None else u[lo:hi],
                            seed=seed, point_offset=st.point_offset + lo, log_r=st.log_r[lo:hi],
                            x_sample=st.x_sample[lo:hi], z=st.z[lo:hi], elbo_acc=st.elbo_acc, workspace=st.workspace)
            free[slot].record(cur)
            core.suffstats(st.x_sample[lo:hi], st.log_r[lo:hi], r_is_log=True, stats=st.stats)
        if st.use_dist:
            torch.distributed.all_reduce(st.red, group=st.pg)
        core.ng_update(st.stats, rho, prior, theta)
        out = dict(elbo_acc=st.elbo_acc)
    if copy_back is not None:
        copy_back[0].copy_(st.log_r, non_blocking=True)
        copy_back[1].copy_(st.x_sample, non_blocking=True)
    # one packed device buffer -> ONE device-to-host copy for the five global parameters (each .to('cpu') is a synchronising
    # copy of its own; at the launch-bound shapes that is most of the end-to-end time)
    flat = torch.cat([t.reshape(-1) for t in theta]).to('cpu')
    theta_h, o = [], 0
    for t in theta:
        theta_h.append(flat[o:o + t.numel()].reshape(t.shape))
        o += t.numel()
    elbo = out['elbo_acc'].to('cpu', non_blocking=False)
    return elbo.numpy(), theta_h
