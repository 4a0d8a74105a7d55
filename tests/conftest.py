import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False) as f:
        return {k: f[k] for k in f.files}


def T(a, dtype=torch.float64, device='cpu'):
    return torch.as_tensor(np.asarray(a), dtype=dtype, device=device)


def regen_noise(g):
    """The draws the shim made for tf.random_normal / tf.multinomial (see make_golden.py)."""
    N, K, D, S, seed = int(g['N']), int(g['K']), int(g['D']), int(g['S']), int(g['seed'])
    noise = np.random.RandomState(seed).standard_normal((N, K, D, S))
    u = np.random.RandomState(seed).random_sample((N, S))
    return noise, u


def regen_decoder(g):
    N, K, S = int(g['N']), int(g['K']), int(g['S'])
    rs = np.random.RandomState(int(g['rec_seed']))
    dobs = int(g['dobs'])
    y = rs.randn(N, dobs)
    decoder = str(g['decoder']) if 'decoder' in g else 'standard'
    if decoder == 'standard':
        return y, (rs.randn(N, K, S, dobs), np.exp(0.3 * rs.randn(N, K, S, dobs))), decoder
    return np.sign(y), (np.zeros((N, K, S, dobs)), rs.randn(N, K, S, dobs)), decoder


SVAE_CASES = ['svae_c1', 'svae_c2', 'svae_init', 'svae_d8', 'svae_d16', 'svae_d32', 'svae_d64', 'svae_d32_overlap',
              'svae_d64_overlap']


def losses_inputs(seed, N, K, S, D, C=4):
    """Same seeded inputs as tests/golden/make_golden.py::losses_inputs."""
    rs = np.random.RandomState(seed)
    y = rs.randn(N, D)
    pred = y[:, None, None, :] + 0.7 * rs.randn(N, K, S, D)
    var = np.exp(0.4 * rs.randn(N, K, S, D))
    logits = 1.5 * rs.randn(N, K, S, D)
    r = rs.dirichlet(0.5 * np.ones(K), N)
    labels = rs.randint(0, C, N)
    return y, pred, var, logits, r, labels


def imputation_stub(N, K, S, D):
    """Same deterministic stand-in for svae.inference as make_golden.py::imputation_stub (numpy in, numpy out)."""
    cnt = [0]

    def imp(y_pert):
        rs = np.random.RandomState(500 + cnt[0])
        cnt[0] += 1
        means = y_pert[:, None, None, :] + 0.3 * rs.randn(N, K, S, D)
        return means, np.exp(0.3 * rs.randn(N, K, S, D)), np.log(rs.dirichlet(np.ones(K), N))
    return imp
