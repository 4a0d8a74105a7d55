"""Two-rank NCCL run of the sharded step on real GPUs (skipped when fewer than 2 devices are visible): points are
sharded contiguously, the packed statistics are all-reduced once, and every rank must end with the theta a single
GPU computes from the full batch — with injected noise AND with the in-kernel Philox stream (keyed by the global
pair index: every rank passes the same seed and its shard's point_offset)."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
from vmp_for_svae_b200 import synthetic
from vmp_for_svae_b200.dist import init_from_env, shard_range
from vmp_for_svae_b200.step import SVAEStep
rank, world, local = init_from_env('nccl')
dev = torch.device('cuda', local)
N, K, D, S = 4096, 12, 16, 1
dt = torch.float32
prior, theta, phi_gmm = synthetic.make_globals(K, D, seed=0, dtype=dt, device=dev)
eta1, eta2d = synthetic.make_encoder_outputs(N, D, synthetic.cluster_centres(phi_gmm), seed=1, dtype=dt, device=dev, spread=0.5)
g = torch.Generator(device=dev).manual_seed(5)
noise = torch.randn(N, K, D, S, generator=g, dtype=dt, device=dev)
u = torch.rand(N, K, generator=g, dtype=dt, device=dev).clamp_(1e-6, 1 - 1e-6)
# full batch on this GPU alone
th_full = [t.clone() for t in theta]
full = SVAEStep(N, K, D, S, dtype=dt, device=dev, use_dist=False)
of = full.step((eta1, eta2d), phi_gmm, th_full, prior, 0.3, noise=noise, u=u)
acc_full = of['elbo_acc'].clone()
# sharded across the two ranks
a, b = shard_range(N, rank, world)
th = [t.clone() for t in theta]
st = SVAEStep(b - a, K, D, S, dtype=dt, device=dev)
assert st.use_dist
o = st.step((eta1[a:b].contiguous(), eta2d[a:b].contiguous()), phi_gmm, th, prior, 0.3,
            noise=noise[a:b].contiguous(), u=u[a:b].contiguous())
torch.cuda.synchronize()
assert torch.equal(o['log_r'], of['log_r'][a:b]) and torch.equal(o['z'], of['z'][a:b])
for x, y in zip(th, th_full):
    torch.testing.assert_close(x, y, rtol=1e-6, atol=1e-6)
torch.testing.assert_close(o['elbo_acc'][:3], acc_full[:3], rtol=1e-9, atol=1e-6)
# the same with in-kernel noise: seed shared by all ranks, stream keyed by the global point index
th_full2 = [t.clone() for t in theta]
of2 = SVAEStep(N, K, D, S, dtype=dt, device=dev, use_dist=False).step((eta1, eta2d), phi_gmm, th_full2, prior, 0.3, seed=99)
th2 = [t.clone() for t in theta]
st2 = SVAEStep(b - a, K, D, S, dtype=dt, device=dev, point_offset=a)
o2 = st2.step((eta1[a:b].contiguous(), eta2d[a:b].contiguous()), phi_gmm, th2, prior, 0.3, seed=99)
torch.cuda.synchronize()
assert torch.equal(o2['log_r'], of2['log_r'][a:b]) and torch.equal(o2['z'], of2['z'][a:b])
assert torch.equal(o2['x_sample'], of2['x_sample'][a:b])
for x, y in zip(th2, th_full2):
    torch.testing.assert_close(x, y, rtol=1e-6, atol=1e-6)
# replicas are bit-identical after the all-reduce
for t in th:
    other = t.clone()
    dist.broadcast(other, src=0)
    assert torch.equal(other, t)
dist.barrier()
dist.destroy_process_group()
print('RANK_OK', rank)
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_two_rank_nccl_step_matches_single_gpu(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % dict(root=ROOT))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', '29541', str(script)]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0 and 'RANK_OK 0' in p.stdout and 'RANK_OK 1' in p.stdout, p.stdout[-3000:]
