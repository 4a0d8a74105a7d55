"""Build libvmp_svae.so in-tree with nvcc for sm_100a (no torch dependency in the library).

One object per source, compiled in parallel; an object is rebuilt only when its source or any header is newer."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libvmp_svae.so')
HEADER = os.path.join(HERE, '..', 'include', 'vmp_svae.h')
SOURCES = ['fast2d_d64.cu', 'fast_d64.cu', 'fast_d64w.cu', 'fast_d32.cu', 'fast_d32w.cu', 'fast_d16.cu', 'fast_d16w.cu', 'fast_pack.cu',
           'prepare.cu', 'local_step.cu', 'local_step_bwd.cu', 'suffstats.cu', 'suffstats_tc.cu', 'mixtures.cu', 'elbo_terms.cu', 'probe.cu']
NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-Xcompiler', '-fPIC', '-Xptxas', '-v', '-Wno-deprecated-gpu-targets']


def _headers_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))] + [HEADER]
    return max(os.path.getmtime(h) for h in hs)


def build(force=False, verbose=False):
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    bdir = os.path.join(HERE, 'build')
    os.makedirs(bdir, exist_ok=True)
    hm = _headers_mtime()
    procs, objs = [], []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        obj = os.path.join(bdir, src.replace('.cu', '.o'))
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(path), hm):
            continue
        cmd = [nvcc] + NVCC_FLAGS + ['-c', path, '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        with open(os.path.join(bdir, src.replace('.cu', '.ptxas.log')), 'w') as f:
            f.write(out)
        if p.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s' % (src, out))
        if verbose:
            print('==== %s\n%s' % (src, out))
    if procs or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        subprocess.check_call([nvcc, '-shared', '-o', LIB] + objs + ['-lcudart'])
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
