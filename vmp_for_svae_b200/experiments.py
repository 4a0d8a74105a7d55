"""Training driver: the BEHAVIOUR of the reference's experiments.py on the B200 path.

One training iteration (experiments.py:208-260):
    phi_enc            = encoder(y)                                  torch MLP (vae.make_encoder, vae.py:75-135)
    x_k, log r, reg    = fused local step (+ selected sample, z)     csrc/local_step*.cu   [autograd.local_step_autograd]
    reconstructions    = decoder(x_k)                                torch MLP (vae.make_decoder, vae.py:138-151)
    neg_rec            = weighted decoder log-likelihood             csrc/elbo_terms.cu    [autograd.decoder_loglike_autograd]
    Adam on -(neg_rec - reg) for encoder, decoder, phi_gmm (+ mu_k, L_k in the SMM variant)
    CVI on theta:  theta <- (1-rho_t) theta + rho_t theta*(x_samples, r),  rho_t = lrcvi * decay^(t/1000)
                   (suffstats.cu + ng_update; the SMM variant updates alpha only, experiments.py:255-256)

The nets are ordinary torch.nn modules (library GEMMs — plumbing around the hot path); everything between them runs in
the hand-written kernels.  Data: pinwheel generator (data.py:216-235); 'auto' needs the UCI csv, which is not shipped,
so `make_dataset('auto-like')` draws a 6-D synthetic stand-in of the same size and scaling.
"""
import math
import time

import numpy as np
import torch

from . import core
from .autograd import decoder_loglike_autograd, local_step_autograd, student_theta_record
from .losses import (bernoulli_logprob, diagonal_gaussian_logprob, generate_missing_data_mask, imputation_losses, purity,
                     weighted_mse)
from .models import svae


# ------------------------------------------------------------------------------------------------ schedule / data
def create_schedule(param_ranges):
    """helpers/scheduling.py:7-36 : cartesian product of the list-valued entries -> list of config dicts."""
    keys, lists = [], []
    for k, v in param_ranges.items():
        if isinstance(v, str) or not hasattr(v, '__iter__'):
            v = [v]
        keys.append(k)
        lists.append(list(v))
    out = [{}]
    for k, vs in zip(keys, lists):
        out = [dict(c, **{k: v}) for c in out for v in vs]
    return out


def make_pinwheel_data(radial_std, tangential_std, num_classes, num_per_class, rate, seed=1):
    """data.py:216-235 (Johnson et al. 2016): spokes of a pinwheel, shuffled; -> (data[N,2], labels[N])."""
    rs = np.random.RandomState(seed)
    rads = np.linspace(0, 2 * np.pi, num_classes, endpoint=False)
    feats = rs.randn(num_classes * num_per_class, 2) * np.array([radial_std, tangential_std])
    feats[:, 0] += 1.0
    labels = np.repeat(np.arange(num_classes), num_per_class)
    ang = rads[labels] + rate * np.exp(feats[:, 0])
    rot = np.stack([np.cos(ang), -np.sin(ang), np.sin(ang), np.cos(ang)]).T.reshape(-1, 2, 2)
    pts = 10.0 * np.einsum('ti,tij->tj', feats, rot)
    perm = rs.permutation(pts.shape[0])
    return pts[perm], labels[perm].astype(np.int64)


def perturb_data(x, noise_ratio=0.1, noise_mean=0.0, noise_stddev=10.0, seed=0):
    """data.py:238-259 : replace a random subset of the points by N(mean, std) noise."""
    rs = np.random.RandomState(seed)
    x = np.array(x, copy=True)
    n = x.shape[0]
    idx = rs.choice(n, int(noise_ratio * n), replace=False)
    x[idx] = noise_mean + noise_stddev * rs.randn(idx.size, x.shape[1])
    return x


def make_dataset(dataset, ratio_tr=0.7, seed_split=0, noise_level=0.1, path_datadir='../datasets'):
    """data.py:9-128 for the array datasets -> (X_tr, lbl_tr, X_te, lbl_te) numpy; labels are class indices.
    'pinwheel' / 'noisy-pinwheel' are generated; 'auto', 'geyser', 'aggregation' read the reference's files from
    `path_datadir` when present (they are not shipped); 'auto-like' is a synthetic 6-D stand-in for 'auto'."""
    import os
    if dataset in ('pinwheel', 'noisy-pinwheel'):
        data, labels = make_pinwheel_data(0.3, 0.05, 5, 200, 0.25)
    elif dataset == 'auto-like':
        rs = np.random.RandomState(7)
        labels = rs.randint(0, 5, 392)
        centres, mix = 2.0 * rs.randn(5, 6), rs.randn(5, 6, 6) * 0.4
        data = centres[labels] + np.einsum('nij,nj->ni', mix[labels], rs.randn(392, 6))
    elif dataset in ('auto', 'geyser', 'aggregation'):
        import pandas as pd
        path = {'auto': 'Auto/auto-mpg.csv', 'geyser': 'geyser', 'aggregation': 'Aggregation.txt'}[dataset]
        path = os.path.join(path_datadir, path)
        if not os.path.exists(path):
            raise FileNotFoundError("dataset '%s' needs %s (not shipped with this repository); use 'auto-like' or "
                                    "'pinwheel' for a self-contained run" % (dataset, path))
        if dataset == 'auto':                                         # data.py:56-75
            raw = pd.read_csv(path, sep=',', header=None).values
            raw = raw[raw[:, 3] != '?']
            cyl = raw[:, 1].astype(np.int64)
            labels = np.select([cyl == 3, cyl == 4, cyl == 5, cyl == 6, cyl == 8], [0, 1, 2, 3, 4])
            data = raw[:, [0, 2, 3, 4, 5, 6]].astype(np.float64)
        elif dataset == 'geyser':                                     # data.py:41-44
            data = pd.read_csv(path, sep=' ', header=None).values[:, [1, 2]].astype(np.float64)
            labels = (data[:, 1] > 20).astype(np.int64)
        else:                                                         # data.py:50-54
            raw = pd.read_csv(path, sep='\t', header=None).values
            data, labels = raw[:, 0:2].astype(np.float64), raw[:, 2].astype(np.int64) - 1
    else:
        raise Exception("Dataset '%s' does not exist." % dataset)
    rs = np.random.RandomState(seed_split)
    perm = rs.permutation(data.shape[0])
    n_te = int(math.ceil((1.0 - ratio_tr) * data.shape[0]))
    te, tr = perm[:n_te], perm[n_te:]
    X_tr, X_te = data[tr], data[te]
    if dataset == 'noisy-pinwheel':
        X_tr = perturb_data(X_tr, noise_ratio=noise_level, seed=seed_split)
    if dataset in ('auto-like', 'auto'):                              # data.py:114-117 : standardise, times 5
        mu, sd = X_tr.mean(0), X_tr.std(0)
        X_tr, X_te = (X_tr - mu) / sd * 5.0, (X_te - mu) / sd * 5.0
    elif dataset not in ('pinwheel', 'noisy-pinwheel'):               # data.py:118-121
        mu, sd = X_tr.mean(0), X_tr.std(0)
        X_tr, X_te = (X_tr - mu) / sd, (X_te - mu) / sd
    return X_tr, labels[tr], X_te, labels[te]


# ------------------------------------------------------------------------------------------------ networks
def rand_partial_isometry(m, n, stddev, seed=0):
    """vae.py:58-72."""
    d = max(m, n)
    return np.linalg.qr(np.random.RandomState(seed).normal(0.0, stddev, size=(d, d)))[0][:m, :n]


class ResNet(torch.nn.Module):
    """vae.make_nnet (vae.py:75-128): tanh MLP + linear shortcut; head 'standard' -> (mean, softplus var),
    'natparam' -> (eta1, -softplus/2), 'bernoulli' -> logits.  Inputs of any leading shape [..., Din]."""

    def __init__(self, in_dim, layerspecs, stddev=0.01, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(int(seed))
        dims = [in_dim] + [u for u, _ in layerspecs[:-1]]
        self.out_dim, self.kind = layerspecs[-1]
        heads = 1 if self.kind == 'bernoulli' else 2
        self.hidden = torch.nn.ModuleList([torch.nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:])])
        self.head = torch.nn.Linear(dims[-1], heads * self.out_dim)
        for lin in list(self.hidden) + [self.head]:                   # vae.py:17-24 : N(0, stddev) weights AND biases
            with torch.no_grad():
                lin.weight.copy_(stddev * torch.randn(lin.weight.shape, generator=g))
                lin.bias.copy_(stddev * torch.randn(lin.bias.shape, generator=g))
        self.W = torch.nn.Parameter(torch.as_tensor(rand_partial_isometry(in_dim, self.out_dim, 1.0, seed), dtype=torch.float32))
        self.b1 = torch.nn.Parameter(torch.zeros(self.out_dim))
        self.b2 = torch.nn.Parameter(torch.zeros(self.out_dim)) if heads == 2 else None

    def forward(self, x):
        lead = x.shape[:-1]
        h = x.reshape(-1, x.shape[-1])
        inp = h
        for lin in self.hidden:
            h = torch.tanh(lin(h))
        o = self.head(h)
        res = inp @ self.W + self.b1
        if self.kind == 'bernoulli':
            logits = (o + res).reshape(*lead, self.out_dim)
            return torch.sigmoid(logits), logits                      # vae.py:146-149
        raw1, raw2 = o[:, :self.out_dim], o[:, self.out_dim:]
        sp, a = torch.nn.functional.softplus, (1.0 if self.kind == 'standard' else -0.5)
        return (raw1 + res).reshape(*lead, self.out_dim), (a * sp(raw2) + a * sp(self.b2)).reshape(*lead, self.out_dim)


# ------------------------------------------------------------------------------------------------ the loop
class SVAETrainer(object):
    """State + one-iteration step of the reference's training graph for config['method'] in
    {'svae-cvi', 'svae-cvi-smm'} (experiments.py:101-260)."""

    def __init__(self, config, obs_dim, device='cuda', nb_samples=10, stddev_init_nn=0.01, decoder_type='standard'):
        self.cfg, self.dev, self.S, self.decoder_type = dict(config), torch.device(device), nb_samples, decoder_type
        K, L, U, seed = config['K'], config['L'], config['U'], config.get('seed', 0)
        self.K, self.L = K, L
        self.smm = 'smm' in config['method']
        self.encoder = ResNet(obs_dim, [(U, 'tanh'), (U, 'tanh'), (L, 'natparam')], stddev_init_nn, seed).to(self.dev)
        self.decoder = ResNet(L, [(U, 'tanh'), (U, 'tanh'), (obs_dim, decoder_type)], stddev_init_nn, seed).to(self.dev)
        prior, theta = svae.init_mm(K, L, seed=seed, param_device=self.dev)
        self.phi_gmm = [torch.nn.Parameter(t.clone()) for t in svae.init_recognition_params(theta, K, seed=seed)]
        params = list(self.encoder.parameters()) + list(self.decoder.parameters()) + self.phi_gmm
        if self.smm:                                                  # experiments.py:154-178
            mu_k, L_k = svae.make_loc_scale_variables(prior)
            self.mu_k, self.L_k = torch.nn.Parameter(mu_k.clone()), torch.nn.Parameter(L_k.clone())
            self.dof = config['DoF'] * torch.ones(K, device=self.dev)
            self.alpha = theta[0].clone()
            self.prior = [prior[0]]
            params += [self.mu_k, self.L_k]
        else:
            self.prior, self.theta = prior, theta
        self.opt = torch.optim.Adam(params, lr=config['lr'], eps=1e-8)   # tf.train.AdamOptimizer defaults
        self.global_step = 0

    def lrcvi(self):
        """experiments.py:143-147 : exponential_decay(lrcvi, step, 1000, decay_rate)."""
        return self.cfg['lrcvi'] * self.cfg.get('decay_rate', 1.0) ** (self.global_step / 1000.0)

    def theta_record(self):
        if self.smm:
            return student_theta_record(self.alpha, self.mu_k, self.L_k, self.dof)
        return core.theta_prepare_gauss(self.theta)

    def forward(self, y, S, seed, train=True):
        eta1, eta2d = self.encoder(y)
        den = core.DEN_STUDENT if self.smm else core.DEN_GAUSS
        x_k, log_r, reg, acc, x_samp, z = local_step_autograd(
            eta1, eta2d, self.phi_gmm[0], self.phi_gmm[1], self.phi_gmm[2], self.theta_record(), S, den_mode=den,
            seed=seed, full=True)
        rec = self.decoder(x_k)
        return rec, x_k, log_r, reg, acc, x_samp, z

    def train_step(self, y):
        """One sess.run(training_step): Adam on the deterministic parameters + CVI on theta.  -> dict of floats
        (device scalars; call .item() on them only when logging)."""
        seed = (self.cfg.get('seed', 0) << 32) + self.global_step
        rec, x_k, log_r, reg, acc, x_samp, z = self.forward(y, self.S, seed)
        neg_rec = decoder_loglike_autograd(y, rec, torch.exp(log_r), self.decoder_type)
        elbo = neg_rec - reg
        self.opt.zero_grad(set_to_none=True)
        (-elbo).backward()
        rho = self.lrcvi()
        # theta update uses the samples / responsibilities of THIS pass (experiments.py:246-256)
        if self.smm:
            stats = core.suffstats(torch.zeros(y.shape[0], 1, device=self.dev), log_r.detach(), r_is_log=True)
            core.ng_update(stats, rho, self.prior, [self.alpha], only_alpha=True)
        else:
            stats = core.suffstats(x_samp, log_r.detach(), r_is_log=True)
            core.ng_update(stats, rho, self.prior, self.theta)
        self.opt.step()
        self.global_step += 1
        return dict(elbo=elbo.detach(), neg_rec=neg_rec.detach(), reg=reg.detach(), bad_pivots=acc[3])

    def train_step_streamed(self, y, tile_points):
        """The same iteration as `train_step`, with x_k_samples STREAMED to the decoder tile by tile (SURVEY 8f row 2;
        svae.py:511 `decoder(x_k_samples)` on the whole [N,K,S,D] tensor, vae.py:138-151): the points are cut into tiles of
        `tile_points`; per tile: encoder -> fused local step (its samples [tile,K,S,D] only) -> decoder -> weighted
        log-likelihood -> backward, which frees the tile's samples and decoder activations before the next tile is produced.
        The ELBO is a sum over points given (phi_gmm, theta), so parameter gradients simply accumulate over the tiles and the
        statistics of the CVI step accumulate in one buffer.  The noise is the batch-level Philox stream (keyed by the global
        point index), i.e. exactly the draws of the untiled call."""
        seed = (self.cfg.get('seed', 0) << 32) + self.global_step
        N, K, L, S = y.shape[0], self.K, self.L, self.S
        den = core.DEN_STUDENT if self.smm else core.DEN_GAUSS
        self.opt.zero_grad(set_to_none=True)
        stats, tot, bad = None, torch.zeros(3, device=self.dev), 0
        for lo in range(0, N, int(tile_points)):
            hi = min(N, lo + int(tile_points))
            yt = y[lo:hi]
            noise, u = core.fill_noise(hi - lo, K, L, S, seed, torch.float32, self.dev, point_offset=lo)
            eta1, eta2d = self.encoder(yt)
            x_k, log_r, reg, acc, x_samp, z = local_step_autograd(
                eta1, eta2d, self.phi_gmm[0], self.phi_gmm[1], self.phi_gmm[2], self.theta_record(), S, den_mode=den,
                noise=noise, u=u, full=True)
            neg_rec = decoder_loglike_autograd(yt, self.decoder(x_k), torch.exp(log_r), self.decoder_type)
            (-(neg_rec - reg)).backward()
            with torch.no_grad():
                tot += torch.stack([neg_rec.detach() - reg.detach(), neg_rec.detach(), reg.detach()])
                xs = torch.zeros(hi - lo, 1, device=self.dev) if self.smm else x_samp
                stats = core.suffstats(xs, log_r.detach(), r_is_log=True, stats=stats)
                bad = bad + acc[3]
            del x_k, log_r, reg, neg_rec, noise
        rho = self.lrcvi()
        if self.smm:
            core.ng_update(stats, rho, self.prior, [self.alpha], only_alpha=True)
        else:
            core.ng_update(stats, rho, self.prior, self.theta)
        self.opt.step()
        self.global_step += 1
        return dict(elbo=tot[0], neg_rec=tot[1], reg=tot[2], bad_pivots=bad)

    @torch.no_grad()
    def evaluate(self, y, labels=None, nb_samples=100, seed=12345):
        """experiments.py:262-304 : test-time inference with S=100 -> mse, log-likelihood, (entropy, purity);
        labels are class indices."""
        rec, x_k, log_r, reg, acc, _, _ = self.forward(y, nb_samples, seed, train=False)
        means, out2 = rec
        out = dict(mse=float(weighted_mse(y, means, torch.exp(log_r))))
        if self.decoder_type == 'standard':
            out['loli'] = float(diagonal_gaussian_logprob(y, means, out2, log_r))
        else:
            out['loli'] = float(bernoulli_logprob(y, out2, log_r))
        if labels is not None:
            onehot = torch.nn.functional.one_hot(labels.long(), int(labels.max()) + 1)
            ent, pur = purity(torch.exp(log_r), onehot)
            out['entropy'], out['purity'] = float(ent), float(pur)
        return out

    @torch.no_grad()
    def imputation(self, y, ratio_missing_data=0.1, nb_samples_pert=20, nb_samples=100, seed=0):
        """experiments.py:361-377 : missing-data imputation through svae.inference -> (imp_mse, imp_logprob)."""
        mask = generate_missing_data_mask(y, ratio_missing_data, seed=seed)
        cnt = [0]

        def impute(y_perturbed):
            cnt[0] += 1
            rec, _, log_r, _, _, _, _ = self.forward(y_perturbed.contiguous(), nb_samples, seed * 7919 + cnt[0], train=False)
            return rec[0], rec[1], log_r
        mse, ll = imputation_losses(y, mask, impute, nb_samples_pert, nb_samples, seed=seed, decoder_type=self.decoder_type)
        return float(mse), float(ll)


class GraphedSVAETrainer(SVAETrainer):
    """The whole training iteration (minibatch gather, encoder, fused local step, decoder, both reverse kernels, Adam,
    statistics, CVI update with the decaying step size) captured ONCE in a CUDA graph and replayed: the reference's
    shapes (64-100 points, K = 10) are launch-bound, ~40 small launches per iteration.  Everything an iteration needs
    is device-resident: minibatch indices, noise and Gumbel uniforms come from torch's graph-safe generator, the CVI
    step size is a device scalar multiplied by decay^(1/1000) each replay, the upstream scalar of the regulariser
    reaches the reverse kernel as a device pointer.  Both variants ('svae-cvi', 'svae-cvi-smm')."""

    def __init__(self, config, y_train, size_minibatch, device='cuda', nb_samples=10, stddev_init_nn=0.01,
                 decoder_type='standard'):
        super().__init__(config, y_train.shape[1], device=device, nb_samples=nb_samples, stddev_init_nn=stddev_init_nn,
                         decoder_type=decoder_type)
        self.y_train, self.M = y_train, int(size_minibatch)
        params = [p for grp in self.opt.param_groups for p in grp['params']]     # incl. mu_k, L_k of the SMM variant
        self.opt = torch.optim.Adam(params, lr=config['lr'], eps=1e-8, capturable=True)
        self.rho = torch.full((1,), float(config['lrcvi']), dtype=torch.float64, device=self.dev)
        self.decay = float(config.get('decay_rate', 1.0)) ** (1.0 / 1000.0)
        self.out = None
        self.graph = None

    def _iteration(self):
        N, K, L, S = self.M, self.K, self.L, self.S
        idx = torch.randint(0, self.y_train.shape[0], (N,), device=self.dev)
        y = self.y_train[idx]
        noise = torch.randn(N, K, L, S, device=self.dev)
        u = torch.rand(N, K, device=self.dev)
        eta1, eta2d = self.encoder(y)
        x_k, log_r, reg, acc, x_samp, z = local_step_autograd(
            eta1, eta2d, self.phi_gmm[0], self.phi_gmm[1], self.phi_gmm[2], self.theta_record(), S,
            den_mode=core.DEN_STUDENT if self.smm else core.DEN_GAUSS, noise=noise, u=u, full=True)
        neg_rec = decoder_loglike_autograd(y, self.decoder(x_k), torch.exp(log_r), self.decoder_type)
        elbo = neg_rec - reg
        (-elbo).backward()
        if self.smm:                                                     # experiments.py:255-256 : alpha only
            stats = core.suffstats(torch.zeros(N, 1, device=self.dev), log_r.detach(), r_is_log=True)
            core.ng_update(stats, self.rho, self.prior, [self.alpha], only_alpha=True)
        else:
            stats = core.suffstats(x_samp, log_r.detach(), r_is_log=True)
            core.ng_update(stats, self.rho, self.prior, self.theta)
        self.opt.step()
        self.rho.mul_(self.decay)                                       # experiments.py:143-147
        return torch.stack([elbo.detach(), neg_rec.detach(), reg.detach(), acc[3].to(elbo.dtype)])

    def capture(self, warmup=3):
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.opt.zero_grad(set_to_none=True)
                self._iteration()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        self.global_step += warmup
        self.graph = torch.cuda.CUDAGraph()
        self.opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.out = self._iteration()                                # recorded, not executed
        return self

    def train_step(self, y=None):
        """One replay = one training iteration on a fresh on-device minibatch.  -> tensor [elbo, neg_rec, reg, bad]."""
        if self.graph is None:
            self.capture()
        self.graph.replay()
        self.global_step += 1
        return self.out


def run_experiment(config, nb_iters=2000, size_minibatch=64, measurement_freq=500, device='cuda', verbose=True,
                   nb_samples=10, nb_samples_te=100, graphed=False):
    """experiments.py:101-478 without TF session plumbing, plots and checkpoints.  -> (trainer, history list)."""
    torch.manual_seed(config.get('seed', 0))
    X_tr, l_tr, X_te, l_te = make_dataset(config['dataset'], noise_level=config.get('noise_level', 0.1))
    dev = torch.device(device)
    y_tr = torch.as_tensor(X_tr, dtype=torch.float32, device=dev)
    y_te = torch.as_tensor(X_te, dtype=torch.float32, device=dev)
    lbl_te = torch.as_tensor(l_te, device=dev)
    if graphed:
        tr = GraphedSVAETrainer(config, y_tr, size_minibatch, device=dev, nb_samples=nb_samples).capture()
    else:
        tr = SVAETrainer(config, y_tr.shape[1], device=dev, nb_samples=nb_samples)
    g = torch.Generator(device='cpu').manual_seed(config.get('seed', 0))
    hist, t0 = [], time.time()
    for i in range(nb_iters):
        if graphed:
            o = tr.train_step()
            out = dict(elbo=o[0], neg_rec=o[1], reg=o[2], bad_pivots=o[3])
        else:
            idx = torch.randint(0, y_tr.shape[0], (size_minibatch,), generator=g).to(dev)   # shuffle_batch stand-in
            out = tr.train_step(y_tr[idx].contiguous())
        if i % measurement_freq == 0 or i == nb_iters - 1 or i == 1:
            ev = tr.evaluate(y_te, lbl_te, nb_samples=nb_samples_te)
            ev.update(iter=i, neg_elbo_normed=-float(out['elbo']) / size_minibatch, sec=time.time() - t0,
                      neg_rec=float(out['neg_rec']), reg=float(out['reg']),
                      lrcvi=float(tr.rho) if graphed else tr.lrcvi(),
                      bad_pivots=float(out['bad_pivots']))
            if i == nb_iters - 1:
                ev['imp_mse'], ev['imp_logprob'] = tr.imputation(y_te, nb_samples=nb_samples_te)
            hist.append(ev)
            if verbose:
                print('Iteration %5d\t%.2fs\t-elbo/M %.4f\tmse_te %.4f\tloli_te %.4f\tpurity %.3f' % (
                    i, ev['sec'], ev['neg_elbo_normed'], ev['mse'], ev.get('loli', float('nan')), ev.get('purity', float('nan'))))
    return tr, hist


if __name__ == '__main__':
    import argparse
    import json
    ap = argparse.ArgumentParser()
    ap.add_argument('--dataset', default='pinwheel')
    ap.add_argument('--method', default='svae-cvi')
    ap.add_argument('--iters', type=int, default=2000)
    ap.add_argument('--out', default=None)
    ap.add_argument('--graphed', action='store_true', help='replay the iteration from a CUDA graph')
    a = ap.parse_args()
    pin = a.dataset != 'auto-like'
    cfg = create_schedule({'dataset': a.dataset, 'method': a.method, 'lr': [0.01 if pin else 0.0003],
                           'lrcvi': [0.1 if pin else 0.2], 'decay_rate': [1.0 if pin else 0.95], 'K': 10,
                           'L': [2 if pin else 6], 'U': 50 if 'smm' not in a.method else 40, 'DoF': 5, 'seed': 0})[0]
    _, hist = run_experiment(cfg, nb_iters=a.iters, size_minibatch=100 if pin else 64, graphed=a.graphed)
    if a.out:
        with open(a.out, 'w') as f:
            json.dump(dict(config=cfg, history=hist), f, indent=1)
