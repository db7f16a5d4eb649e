"""Synthetic workloads of the shapes BASELINE.json names: peptide token batches and a
stand-in for a fitted Q(z) with two z-space attribute classifiers.  Used by bench.py and the examples;
there is no network for the real datasets or checkpoints.
"""
import numpy as np
import torch

PAD_IDX, START_IDX, EOS_IDX = 1, 2, 3          # models/mutils.py:5-8
Z_DIM = 100


def synthetic_tokens(batch, n_vocab, seed=1238, max_len=25):
    """Rows `[<start>, aa * len, <eos>, <pad>...]`, len ~ U{5..23}, amino-acid ids U{4..V-1}; int64 [B, max_len]
    (the layout torchtext's Field(init_token, eos_token, fix_length) produces, data_processing/dataset.py:242-244)."""
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(5, max_len - 1, (batch,), generator=g)
    body = torch.randint(4, n_vocab, (batch, max_len), generator=g)
    pos = torch.arange(max_len).unsqueeze(0)
    n = lens.unsqueeze(1)
    toks = torch.full((batch, max_len), PAD_IDX, dtype=torch.int64)
    toks = torch.where((pos >= 1) & (pos <= n), torch.roll(body, 1, dims=1), toks)
    toks = torch.where(pos == n + 1, torch.full_like(toks, EOS_IDX), toks)
    toks[:, 0] = START_IDX
    return toks


def synthetic_class_setup(seed=1238, n_comp=100):
    """(weights, means, diag covariances, classifiers): a mixture of `n_comp` diagonal Gaussians in the range a
    fitted mogQ has, and two float32 logistic classifiers (amp target 1, tox target 0) giving ~25-30 % acceptance."""
    rs = np.random.RandomState(seed)
    w = rs.dirichlet(np.ones(n_comp) * 5.0)
    means = 0.8 * rs.randn(n_comp, Z_DIM) * 0.5
    covs = rs.uniform(0.3, 0.7, (n_comp, Z_DIM))
    clfs = [('amp', (rs.randn(Z_DIM) * 0.12).astype(np.float32), np.float32(0.3), 1),
            ('tox', (rs.randn(Z_DIM) * 0.12).astype(np.float32), np.float32(-0.3), 0)]
    return w, means, covs, clfs
