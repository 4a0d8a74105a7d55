import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ.get('GRAFT_REPO_ROOT', '/root/repo'))
from vmp_for_svae_b200 import dist as vdist, core
from vmp_for_svae_b200.mixture_step import MixtureSweep
from vmp_for_svae_b200.models import smm
rank, world, local = vdist.init_from_env('nccl')
torch.cuda.set_device(local); dev = torch.device('cuda', local)
K, D, N = 32, 8, 200000
g = torch.Generator(device=dev).manual_seed(100 + rank)
x = torch.randn(N, D, generator=g, device=dev)
e = -torch.log(torch.rand(N, K, generator=g, device=dev)); r = (e / e.sum(1, keepdim=True)).contiguous(); u = torch.ones_like(r)
sw = MixtureSweep(K, D, smm._prior_standard(K, D, 0, torch.float32, dev), kappa_k=torch.full((K,), 5.0, device=dev), device=dev)
for it in range(6):
    out = sw.sweep(x, r, u)
    torch.cuda.synchronize()
    for name, t in (('stats', sw.stats[0]), ('alpha_k', out['alpha_k']), ('C_k', out['C_k']), ('rec', out['rec'])):
        ts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(ts, t.contiguous())
        if rank == 0:
            d = max(float((a.double() - ts[0].double()).abs().max()) for a in ts)
            same = all(torch.equal(a, ts[0]) for a in ts)
            print(it, name, 'identical' if same else 'DIFF max %.3e' % d, 'nan' if any(torch.isnan(a.double()).any() for a in ts) else '')
dist.barrier(); dist.destroy_process_group()
