"""Sharded VB-EM sweeps of the standalone mixtures (BASELINE config C3): points shard contiguously across ranks, the state
(r[, u]) of a shard lives on its rank, the Dirichlet+NIW posterior is replicated, and the ONLY exchange per sweep is one
all-reduce(sum) of the packed K*(D^2+D+2) double statistics between the statistics phase and the K-sized M-step — the same
exchange as the SVAE step (step.py).  Single process: `models.gmm.fit` / `models.smm.fit` (vmp_mixture_fit) do the same."""
import torch

from . import core, _lib


class MixtureSweep(object):
    def __init__(self, K, D, prior_std, kappa_k=None, dtype=torch.float32, device='cuda', process_group=None, use_dist=None):
        self.K, self.D, self.dtype, self.device = int(K), int(D), dtype, torch.device(device)
        self.prior, self.kappa = [t.contiguous() for t in prior_std], kappa_k
        self.is_smm = kappa_k is not None
        slen = _lib.record_lens(D)[2]
        self.stats = [torch.zeros(K, slen, dtype=torch.float64, device=self.device) for _ in range(2)]
        self.fast = dtype == torch.float32 and D <= 8 and K <= 32
        self.out = None
        self.pg = process_group
        if use_dist is None:
            use_dist = torch.distributed.is_available() and torch.distributed.is_initialized() and \
                torch.distributed.get_world_size(process_group) > 1
        self.use_dist = bool(use_dist)

    def _mstep(self, cur, nxt):
        if self.use_dist:
            torch.distributed.all_reduce(cur, group=self.pg)
        self.out = core.mixture_prepare(cur, self.prior, self.is_smm, kappa_k=self.kappa, stats_next=nxt, out=self.out,
                                        want_general=not self.fast)
        return self.out

    def fit(self, x, r, u=None, n_sweeps=1):
        """n_sweeps x (M-step of the current state -> e-step), state updated in place; returns the last M-step's outputs."""
        cur, nxt = self.stats
        cur.zero_()
        core.suffstats(x, r, u_nk=u if self.is_smm else None, stats=cur)
        for s in range(n_sweeps):
            last = s + 1 == n_sweeps
            o = self._mstep(cur, None if last else nxt)
            if self.fast:
                core.mixture_estep_fused(x, o['rec'], self.is_smm, r=r, u=u, stats_next=None if last else nxt, write_state=last)
            else:
                core.mixture_estep(x, o['alpha_k'], o['beta_k'], o['m_k'], o['P_k'], o['v_k'], kappa_k=self.kappa, r=r, u_out=u)
                if not last:
                    core.suffstats(x, r, u_nk=u if self.is_smm else None, stats=nxt)
            cur, nxt = nxt, cur
        return o

    def sweep(self, x, r, u=None):
        """one gmm.inference / smm.inference sweep on this rank's shard (state in, state out)"""
        return self.fit(x, r, u, 1)
