"""CPU baselines for bench.py: the reference's hot path timed on the host cores.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/__init__.py) -- never a fallback.

kind "reference": the UNMODIFIED reference imported from /root/reference (build container only),
                  driven through its own `train_vae.train_vae` / `mogQ.rejection_sample` /
                  `generate_sentences(sample_mode='beam')`.
kind "port":      where the reference tree is absent (the GPU box): the same computation expressed
                  with the same library calls the reference makes (nn.Embedding, nn.GRU, nn.Linear,
                  nn.Dropout, F.cross_entropy, the losses.py formulas, torch.optim.Adam with the
                  duplicated embedding, clip_grad_norm_; numpy/sklearn-equivalent GMM sampling and
                  logistic scoring), with all host threads.  The one deliberate difference: the
                  full-kernel MMD Gram matrices are evaluated in row chunks (identical arithmetic per
                  element; the reference's [B,B,100] broadcast needs 3 x 6.7 GB at B=4096).
"""
import math
import os
import time
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import class_sampling as oc
from . import refharness as rh
from .wae import (DEC_H, EMB_DIM, ENC_H, MAX_SEQ_LEN, MMD_SIGMA, PAD_IDX, RF_DIM, Z_DIM, UNK_IDX, anneal_beta,
                  synthetic_tokens)


def host_threads():
    n = os.cpu_count() or 1
    torch.set_num_threads(n)
    return n


class PortModel(nn.Module):
    """Same layers, same construction order as RNN_VAE.__init__ (models/model.py:47-67)."""
    def __init__(self, n_vocab):
        super().__init__()
        self.word_emb = nn.Embedding(n_vocab, EMB_DIM, PAD_IDX)
        self.enc_rnn = nn.GRU(EMB_DIM, ENC_H, num_layers=1, bidirectional=True, batch_first=True)
        self.q_mu = nn.Linear(2 * ENC_H, Z_DIM)
        self.q_logvar = nn.Linear(2 * ENC_H, Z_DIM)
        self.dec_rnn = nn.GRU(EMB_DIM + DEC_H, DEC_H, batch_first=True)
        self.fc = nn.Linear(DEC_H, n_vocab)

    def vae_params(self):
        ps = [self.word_emb.weight] + list(self.enc_rnn.parameters()) + list(self.q_mu.parameters()) + \
            list(self.q_logvar.parameters()) + [self.word_emb.weight] + list(self.dec_rnn.parameters()) + \
            list(self.fc.parameters())
        return ps                                   # embedding twice, like models/model.py:88-94


def _gram_sum_chunked(x, y, sigma, chunk=512):
    tot = x.new_zeros(())
    diag = []
    for s in range(0, x.shape[0], chunk):
        k = torch.exp(-((x[s:s + chunk, None, :] - y[None, :, :]) ** 2).sum(2) / sigma ** 2)
        tot = tot + k.sum()
        idx = torch.arange(s, min(s + chunk, x.shape[0]))
        diag.append(k[idx - s, idx])
    return tot, torch.cat(diag)


class PortWaeTrainer:
    """One iteration = train_vae.py:24-42 (+ the nine .item() reads of :44-53)."""
    def __init__(self, n_vocab=24, seed=1238):
        torch.manual_seed(seed)
        np.random.seed(seed)
        self.m = PortModel(n_vocab)
        self.opt = torch.optim.Adam(self.m.vae_params(), lr=1e-3)
        self.rf = None
        self.it = 0

    def step(self, tokens):
        m = self.m
        B, L = tokens.shape
        beta = anneal_beta(self.it)
        _, h = m.enc_rnn(m.word_emb(tokens))
        h = torch.cat((h[-2], h[-1]), 1)
        mu, logvar = m.q_mu(h), m.q_logvar(h)
        z = mu + torch.exp(logvar / 2) * torch.randn(B, Z_DIM)
        c = torch.from_numpy(np.random.multinomial(1, [0.5, 0.5], B).astype('float32'))
        data = tokens.clone()
        data[torch.from_numpy(np.random.binomial(1, p=0.3, size=(B, L)).astype('uint8')).bool()] = UNK_IDX
        zc = torch.cat([z, c], 1)
        x = torch.cat([m.word_emb(data), zc.unsqueeze(1).expand(-1, L, -1)], 2)
        out, _ = m.dec_rnn(x, zc.unsqueeze(0))
        logits = m.fc(F.dropout(out, 0.3, True))
        tgt = torch.cat([tokens[:, 1:], torch.full((B, 1), PAD_IDX, dtype=torch.long)], 1)
        recon = F.cross_entropy(logits.view(-1, logits.size(2)), tgt.view(-1), reduction='mean', ignore_index=PAD_IDX)
        kl = torch.mean(0.5 * torch.sum(logvar.exp() + mu ** 2 - 1 - logvar, 1))
        with torch.no_grad():                                   # log-only under z_regu_loss='mmdrf'
            zd, zp = z.detach(), torch.randn(B, Z_DIM)
            s11, d11 = _gram_sum_chunked(zd, zd, MMD_SIGMA)
            s22, d22 = _gram_sum_chunked(zp, zp, MMD_SIGMA)
            s12, d12 = _gram_sum_chunked(zd, zp, MMD_SIGMA)
            mmd = ((s11 + s22 - 2 * s12) - B * (d11 + d22 - 2 * d12).sum()) / (B * (B - 1))
        zp2 = torch.randn(B, Z_DIM)
        if self.rf is None:
            self.rf = (torch.randn(Z_DIM, RF_DIM), math.pi * 2 * torch.rand(RF_DIM))
        w, b = self.rf
        f1 = (torch.cos(z @ w / MMD_SIGMA + b) * (2. / RF_DIM) ** 0.5).mean(0)
        f2 = (torch.cos(zp2 @ w / MMD_SIGMA + b) * (2. / RF_DIM) ** 0.5).mean(0)
        mmdrf = ((f1 - f2) ** 2).sum()
        l1 = logvar.abs().sum(1).mean(0)
        klp = torch.mean(0.5 * torch.sum(logvar.exp() - 1 - logvar, 1))
        loss = recon + beta * mmdrf + 0.0 * l1 + 1e-3 * klp
        self.opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.m.vae_params(), 5.0)
        self.opt.step()
        vals = [mu.data.abs().mean().item(), logvar.data.mean().item(), l1.item(), klp.item(), loss.item(),
                recon.item(), kl.item(), mmd.item(), mmdrf.item()]
        self.it += 1
        return vals


class ReferenceWaeTrainer:
    """The live reference: its own RNN_VAE + train_vae.train_vae, one call per timed block."""
    def __init__(self, n_vocab=24, seed=1238):
        self.ref = rh.load_reference()
        self.model = rh.build_model(n_vocab, seed)
        self.cfgv = self.ref.cfg.vae
        rh.reset_rf_cache()

    def run(self, tokens, n_iters):
        import contextlib
        import io
        cfgv = self.cfgv
        cfgv.s_iter, cfgv.n_iter = 1, n_iters - 1          # it = 1 .. n_iters : no logging iteration (it % 500 != 0)
        cfgv.cheaplog_every, cfgv.expsvlog_every = 10 ** 9, 10 ** 9
        ds = rh.DatasetShim([tokens])
        with contextlib.redirect_stdout(io.StringIO()):
            self.ref.train_vae.train_vae(cfgv, self.model, ds)


def time_wae_cpu(batch, n_vocab=24, steps=2, warmup=1, budget_s=150.0):
    """-> dict(seq_per_s, ms_per_step, kind, cores, steps_timed).  Stops early when `budget_s` is spent."""
    cores = host_threads()
    tokens = synthetic_tokens(batch, n_vocab, seed=1238)
    if rh.reference_available():
        kind = 'reference'
        tr = ReferenceWaeTrainer(n_vocab)
        step = lambda: tr.run(tokens, 1)
    else:
        kind = 'port'
        tr = PortWaeTrainer(n_vocab)
        step = lambda: tr.step(tokens)
    t_start = time.perf_counter()
    for _ in range(warmup):
        step()
        if time.perf_counter() - t_start > budget_s / 2:
            break
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s:
            break
    ms = 1e3 * float(np.mean(times))
    return {'seq_per_s': batch / (ms / 1e3), 'ms_per_step': ms, 'kind': kind, 'cores': cores, 'steps_timed': len(times)}


def synthetic_class_setup(seed=1238, n_comp=100):
    """Deterministic stand-in for a fitted Q and two z-space classifiers (no sklearn fit needed on the
    GPU box): mixture of `n_comp` diagonal Gaussians in the range a fitted mogQ has, float32 logistic
    classifiers giving ~25 % acceptance."""
    rs = np.random.RandomState(seed)
    w = rs.dirichlet(np.ones(n_comp) * 5.0)
    means = 0.8 * rs.randn(n_comp, Z_DIM) * 0.5
    covs = rs.uniform(0.3, 0.7, (n_comp, Z_DIM))
    clfs = [('amp', (rs.randn(Z_DIM) * 0.12).astype(np.float32), np.float32(0.3), 1),
            ('tox', (rs.randn(Z_DIM) * 0.12).astype(np.float32), np.float32(-0.3), 0)]
    return w, means, covs, clfs


def time_class_cpu(n_draws=1_000_000, repeats=1):
    """rejection_sample(n) on the host: sklearn-equivalent GMM draw + LR scoring + accept.  When the
    reference tree is present its own RejSampleBase.rejection_sample is what runs."""
    cores = host_threads()
    w, means, covs, clfs = synthetic_class_setup()
    kind = 'port'
    if rh.reference_available():
        import sklearn.mixture
        kind = 'reference'
        dm = rh.load_reference().density_modeling
        mog = sklearn.mixture.GaussianMixture(n_components=len(w), covariance_type='diag')
        mog.weights_, mog.means_, mog.covariances_ = w, means, covs
        mog.precisions_cholesky_ = 1.0 / np.sqrt(covs)
        Q = dm.mogQ.__new__(dm.mogQ)
        Q.mog = mog
        mk = lambda c: types.SimpleNamespace(predict_proba=lambda x, c=c: np.stack(
            [1 - oc.lr_target_proba(x, c[1], c[2], 1), oc.lr_target_proba(x, c[1], c[2], 1)], 1))
        Q.init_attr_classifiers({c[0]: mk(c) for c in clfs}, {c[0]: c[3] for c in clfs})
        run = lambda: Q.rejection_sample(n_draws)[2]
    else:
        def run():
            rs = np.random
            z, _ = oc.gmm_sample(w, means, covs, n_draws, rs)
            u = rs.uniform(size=n_draws)
            return oc.rejection_accept(z, u, clfs)[1]
    np.random.seed(1238)
    best, acc = None, 0.0
    for _ in range(repeats):
        t0 = time.perf_counter()
        a = run()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        acc = float(np.mean(a))
    return {'draws_per_s': n_draws / best, 'accepted_per_s': acc * n_draws / best, 'accept_rate': acc, 'seconds': best,
            'kind': kind, 'cores': cores, 'n_draws': n_draws}


def time_beam_cpu(params, n=192):
    """Beam decode (beam 5, n_best 3) of n latent points on the host: the reference's own generate_sentences when its
    tree is present (Python loop per sample and step, models/model.py:258-328), else the oracle restatement."""
    import torch
    from . import decode as od
    g = torch.Generator().manual_seed(5)
    z = torch.randn(n, 100, generator=g)
    c = torch.eye(2)[torch.randint(0, 2, (n,), generator=g)]
    kind = 'port'
    if rh.reference_available():
        kind = 'reference'
        model = rh.build_model(params['word_emb.weight'].shape[0])
        model.load_state_dict({k: v for k, v in params.items() if k in model.state_dict()}, strict=False)
        model.eval()
        run = lambda: model.generate_sentences(n, z, c, sample_mode='beam', beam_size=5)
    else:
        run = lambda: od.beam_decode(params, z, c)
    t0 = time.perf_counter()
    run()
    dt = time.perf_counter() - t0
    return {'seq_per_s': n / dt, 'seconds': dt, 'n': n, 'kind': kind, 'cores': host_threads()}
