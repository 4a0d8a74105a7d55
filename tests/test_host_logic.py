"""CPU tests: the C-ABI library loads and exports every symbol include/vmp_svae.h declares, host-side argument
checking, and the multi-rank logic (sharding + packed all-reduce) over gloo with world_size 2."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    from vmp_for_svae_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, 'include', 'vmp_svae.h')).read()
    declared = sorted(set(re.findall(r'^(?:int|size_t)\s+(vmp_\w+)\s*\(', header, flags=re.M)))
    assert len(declared) >= 28
    for name in declared:
        assert hasattr(lib, name), 'libvmp_svae.so does not export %s' % name
    assert sorted(_lib.EXPORTS) == declared, 'ctypes table and header disagree'
    assert lib.vmp_version() >= 100
    for D in (1, 2, 6, 64):
        assert _lib.record_lens(D) == (D * D + 2 * D + 4, D * D + D + 4, D * D + D + 2)


def test_no_cpu_path():
    """The product fails loudly on CPU tensors instead of falling back."""
    from vmp_for_svae_b200 import _lib
    from vmp_for_svae_b200.models import svae
    with pytest.raises(RuntimeError):
        _lib.ptr(torch.zeros(3))
    phi_gmm = (torch.zeros(2, 3), torch.eye(3).repeat(2, 1, 1), torch.zeros(2))
    with pytest.raises(RuntimeError):
        svae.e_step((torch.zeros(4, 3), -torch.ones(4, 3)), phi_gmm, 1)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'vmp_for_svae_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f
                assert 'from oracle' not in src and 'import oracle' not in src, f


def test_product_has_no_runtime_switches():
    """north_star: no multi-backend dispatch — nothing in the product library selects a code path from the environment."""
    csrc = os.path.join(ROOT, 'vmp_for_svae_b200', 'csrc')
    for f in os.listdir(csrc):
        assert 'getenv' not in open(os.path.join(csrc, f)).read(), f


def test_uniform_mapping_never_hits_0_or_1(tmp_path):
    """common.cuh::u32_to_unit on the host (nvcc host compile): the extreme counters map strictly inside (0,1) and the
    mapping is exact in fp32 (the 24-bit form rounded 0xFFFFFFFF to 1.0 -> Gumbel = +inf)."""
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    if not os.path.exists(nvcc):
        pytest.skip('nvcc not available')
    src = tmp_path / 'u.cu'
    src.write_text('#include <cstdio>\n#include "%s"\nint main(){ unsigned v[5]={0u,1u,0x7fffffffu,0xfffffe00u,0xffffffffu};'
                   'for(int i=0;i<5;++i) printf("%%.17g\\n",(double)vmp::u32_to_unit(v[i])); return 0; }\n'
                   % os.path.join(ROOT, 'vmp_for_svae_b200', 'csrc', 'common.cuh'))
    exe = tmp_path / 'u'
    subprocess.check_call([nvcc, '-std=c++17', '-o', str(exe), str(src)], stderr=subprocess.DEVNULL)
    vals = [float(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    assert vals[0] == 0.5 / 8388608 and vals[1] == vals[0]
    assert vals[4] == (8388607 + 0.5) / 8388608 and vals[4] < 1.0 and vals[0] > 0.0
    assert all(0.0 < v < 1.0 for v in vals) and vals == sorted(vals)
    assert np.isfinite(-np.log(-np.log(np.float32(vals[4])))) and np.isfinite(-np.log(-np.log(np.float32(vals[0]))))


def test_engine_lane_numbering_and_spare_uniform(tmp_path):
    """Host compile of the engines' lane numbering (local_step_fast.cuh: FastGeom::GLMASK, bits_deposit / bits_extract) and of
    the spare-bit Gumbel uniform (common.cuh).  The measured shared-memory rule (profiles/r2_smem_lane_bits.md) wants neither
    the group-lane bits nor the pair bits to contain BOTH b0 and b1; the numbering must be a bijection of the 32 lanes; the
    source lane of a group broadcast keeps the pair bits; the spare uniform lies strictly inside (0,1)."""
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    if not os.path.exists(nvcc):
        pytest.skip('nvcc not available')
    csrc = os.path.join(ROOT, 'vmp_for_svae_b200', 'csrc')
    src = tmp_path / 'g.cu'
    src.write_text(r"""
#include <cstdio>
#define VMP_FAST_IMPL
#include "%s/local_step_fast.cuh"
using namespace vmp;
template <int D, int BS> int check() {
    constexpr unsigned M = FastGeom<D, BS>::GLMASK, P = ~M & 31u;
    int bad = 0;
    bad += __builtin_popcount(M) != __builtin_ctz((unsigned)BS);
    bad += (M & 3u) == 3u;                       // gl on both b0 and b1: 4-wavefront row reads
    bad += (P & 3u) == 3u;                       // pair on both b0 and b1: 4-wavefront column broadcasts
    bool seen[32] = {};
    for (unsigned lane = 0; lane < 32; ++lane) {
        const unsigned gl = bits_extract(lane, M), pr = bits_extract(lane, P);
        bad += gl >= (unsigned)BS || pr >= 32u / BS;
        const unsigned id = pr * BS + gl;
        bad += seen[id];
        seen[id] = true;
        bad += (bits_deposit(gl, M) | bits_deposit(pr, P)) != lane;
        for (unsigned l = 0; l < (unsigned)BS; ++l) {                      // source lane of group_bcast(v, l)
            const unsigned src = (lane & P) | bits_deposit(l, M);
            bad += bits_extract(src, P) != pr || bits_extract(src, M) != l;
        }
    }
    return bad;
}
int main() {
    printf("%%d %%d %%d %%d\n", check<64, 16>(), check<32, 8>(), check<32, 4>(), check<16, 4>());
    const uint4 lo = make_uint4(0u, 0u, 0u, 0u), hi = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    const uint4 mix = make_uint4(0x1ffu, 0u, 0u, 0u);
    printf("%%.17g %%.17g %%.17g\n", (double)philox_spare_uniform(lo), (double)philox_spare_uniform(hi), (double)philox_spare_uniform(mix));
    return 0;
}
""" % csrc)
    exe = tmp_path / 'g'
    subprocess.check_call([nvcc, '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', str(exe), str(src)],
                          stderr=subprocess.DEVNULL)
    out = subprocess.check_output([str(exe)], text=True).split()
    assert [int(v) for v in out[:4]] == [0, 0, 0, 0]
    u = [float(v) for v in out[4:]]
    assert u[0] == 0.5 / 8388608 and u[1] == (8388607 + 0.5) / 8388608 and 0.0 < u[0] < u[2] < u[1] < 1.0
    assert u[2] == (0x1ff * 2 ** 14 + 0.5) / 8388608           # the 9 low bits of word x are the top bits of the uniform


def test_shard_range_partitions():
    from vmp_for_svae_b200.dist import shard_range
    for n in (0, 1, 7, 100, 1 << 20):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
from vmp_for_svae_b200.dist import init_from_env, shard_range, allreduce_packed, max_over_ranks
from oracle import svae_port
rank, world, _ = init_from_env(backend='gloo')
assert world == 2
rs = np.random.RandomState(0)
N, K, D = 101, 4, 3
x = torch.as_tensor(rs.randn(N, D)); r = torch.as_tensor(rs.dirichlet(np.ones(K), N))
prior, theta = svae_port.init_mm(K, D, uniform=torch.as_tensor(rs.rand(K, D)))
a, b = shard_range(N, rank, world)
xs, rsh = x[a:b], r[a:b]
slen = D * D + D + 2
buf = torch.zeros(K * slen + 4, dtype=torch.float64)
st = buf[:K * slen].view(K, slen)
st[:, 0] = rsh.sum(0); st[:, 1] = rsh.sum(0)
st[:, 2:2 + D] = torch.einsum('nk,nd->kd', rsh, xs)
st[:, 2 + D:] = torch.einsum('nk,nd,ne->kde', rsh, xs, xs).reshape(K, D * D)
buf[K * slen] = float(b - a)
allreduce_packed(buf)
assert buf[K * slen] == N
# identical natural-gradient update on every rank == full-batch oracle m_step
star = svae_port.m_step(prior, x, r)
Nk = st[:, 0]
mine = [prior[0] + Nk, prior[1] + st[:, 2 + D:].view(K, D, D), prior[2] + st[:, 2:2 + D], prior[3] + Nk, prior[4] + Nk + 1]
for p, q in zip(mine, star):
    assert torch.allclose(p, q, rtol=1e-12, atol=1e-12)
assert max_over_ranks(float(rank), 'cpu') == 1.0
dist.barrier()
print('RANK_OK', rank)
'''


def test_two_rank_sharded_statistics_gloo(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % dict(root=ROOT))
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29533', WORLD_SIZE='2', OMP_NUM_THREADS='1')
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and 'RANK_OK %d' % r in o, o
