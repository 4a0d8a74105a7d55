"""Build libvmp_svae.so in-tree with nvcc for sm_100a (no torch dependency in the library).

One object per source, compiled in parallel; an object is rebuilt only when its source or any header is newer."""
import hashlib
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libvmp_svae.so')
HEADER = os.path.join(HERE, '..', 'include', 'vmp_svae.h')
SOURCES = ['fast_d64.cu', 'fast_d32.cu', 'fast_d16.cu', 'fast_pack.cu',
           'prepare.cu', 'local_step.cu', 'small_step.cu', 'local_step_bwd.cu', 'suffstats.cu', 'suffstats_tc.cu', 'suffstats_mma.cu', 'mixtures.cu', 'mixture_sweep.cu', 'elbo_terms.cu', 'probe.cu']
NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-Xcompiler', '-fPIC', '-Xptxas', '-v', '-Wno-deprecated-gpu-targets']


def _digest(paths, extra=''):
    h = hashlib.sha1(extra.encode())
    for p in paths:
        with open(p, 'rb') as f:
            h.update(f.read())
    return h.hexdigest()


def _headers():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))) + [HEADER]


def _nvcc():
    return os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')


def nvcc_available():
    return os.path.exists(_nvcc()) and os.access(_nvcc(), os.X_OK)


def build(force=False, verbose=False):
    """Incremental by CONTENT (sha1 of the source, every header and the flags, kept in build/manifest.json): file
    times do not survive the copy to the GPU box, and an edited .cu / .cuh must never leave a stale binary behind."""
    nvcc = _nvcc()
    bdir = os.path.join(HERE, 'build')
    os.makedirs(bdir, exist_ok=True)
    mpath = os.path.join(bdir, 'manifest.json')
    try:
        with open(mpath) as f:
            manifest = json.load(f)
    except (OSError, ValueError):
        manifest = {}
    hdig = _digest(_headers(), ' '.join(NVCC_FLAGS))
    procs, objs, want = [], [], {}
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        obj = os.path.join(bdir, src.replace('.cu', '.o'))
        objs.append(obj)
        want[src] = _digest([path], hdig)
        if not force and os.path.exists(obj) and manifest.get(src) == want[src]:
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
    want['__sources__'] = SOURCES
    if procs or not os.path.exists(LIB) or manifest.get('__sources__') != SOURCES:
        subprocess.check_call([nvcc, '-shared', '-o', LIB] + objs + ['-lcudart'])
    if want != manifest:
        with open(mpath, 'w') as f:
            json.dump(want, f, indent=1)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
