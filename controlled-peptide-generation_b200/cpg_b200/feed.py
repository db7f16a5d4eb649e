"""On-device data feed with the reference loader's protocol (data_processing/dataset.py): `next_batch(name).text`
is an int64 [B, max_seq_len] batch, `idx2sentence(s)` maps ids back to words.  The whole tokenised, padded dataset is
resident in HBM (25 bytes per example); a batch is ONE kernel launch: Philox uniforms -> binary search in the cumulative
sampling weights (the WeightedRandomIterator's multinomial with replacement, :60-77) -> gather of the token rows.
At millions of sequences per second the torchtext-style Python loader (one Example object per row) is the bottleneck.

Tokenisation / padding / sample weights are one-time host work and follow the reference:
  * Field(init_token='<start>', eos_token='<eos>', fix_length=L, batch_first=True) (:242-244): rows are
    [<start>] + tokens[:L-2] + [<eos>] + <pad>...   (torchtext 0.3.1 counts init/eos inside fix_length and truncates)
  * df_add_sample_weights (:183-201): weight = max over matching column specifiers of their factor (default 1),
    normalised to sum 1.
"""
import types

import numpy as np
import torch

from . import _lib
from ._lib import check, context, lib, ptr, stream_ptr

SPECIALS = ('<unk>', '<pad>', '<start>', '<eos>')          # ids 0..3, models/mutils.py:5-8
UNK_IDX, PAD_IDX, START_IDX, EOS_IDX = 0, 1, 2, 3


def build_vocab(sequences, fixed=None):
    """itos list: the four specials, then the remaining words by decreasing frequency (ties alphabetical) like
    torchtext's Vocab; `fixed` (a list of words) pins the order instead."""
    if fixed is not None:
        return list(fixed)
    from collections import Counter
    cnt = Counter(w for s in sequences for w in str.split(s))
    words = sorted(cnt, key=lambda w: (-cnt[w], w))
    return list(SPECIALS) + [w for w in words if w not in SPECIALS]


def tokenize_and_pad(sequences, itos, max_seq_len=25):
    """List of space-separated strings -> uint8 [N, max_seq_len] rows in the Field layout described above."""
    stoi = {w: i for i, w in enumerate(itos)}
    out = np.full((len(sequences), max_seq_len), PAD_IDX, dtype=np.uint8)
    for r, s in enumerate(sequences):
        ids = [stoi.get(w, UNK_IDX) for w in str.split(s)][:max_seq_len - 2]
        row = [START_IDX] + ids + [EOS_IDX]
        out[r, :len(row)] = row
    return out


def sample_weights(n, masks_and_factors=()):
    """df_add_sample_weights with sample_prob_factors: masks_and_factors = [(bool mask [n], factor), ...]."""
    w = np.ones(n, dtype=np.float64)
    for mask, factor in masks_and_factors:
        m = np.asarray(mask, dtype=bool) & (w < factor)
        w[m] = factor
    return w / w.sum()


class DeviceDataFeed:
    """Drop-in for the parts of AttributeDataLoader the training loop touches (train_vae.py:24-25,55-57)."""

    def __init__(self, token_rows, weights=None, itos=None, mbsize=32, device=None, seed=1238):
        self.device = _lib.tensor_device(device)
        rows = np.ascontiguousarray(token_rows, dtype=np.uint8)
        self.n, self.L = rows.shape
        w = np.full(self.n, 1.0 / self.n) if weights is None else np.asarray(weights, dtype=np.float64)
        assert w.shape == (self.n,) and (w >= 0).all() and w.sum() > 0
        self.tokens = torch.from_numpy(rows).to(self.device)
        self.cdf = torch.from_numpy(np.cumsum(w / w.sum())).to(self.device)
        self.itos = list(itos) if itos is not None else None
        self.mbsize, self.seed, self.step = int(mbsize), int(seed), 0
        self.last_index = None

    def next_batch(self, iterator_name='train_vae', want_index=False):
        out = torch.empty(self.mbsize, self.L, dtype=torch.int64, device=self.device)
        idx = torch.empty(self.mbsize, dtype=torch.int64, device=self.device) if want_index else None
        check(lib().cpg_feed_batch(context(self.device), stream_ptr(), ptr(self.tokens, torch.uint8), ptr(self.cdf, torch.float64),
                                   self.n, self.L, self.seed, self.step, self.mbsize, ptr(out), ptr(idx)), 'cpg_feed_batch')
        self.step += 1
        self.last_index = idx
        return types.SimpleNamespace(text=out)

    def idx2sentence(self, idxs, print_special_tokens=True):
        ids = [int(i) for i in idxs]
        if not print_special_tokens:
            ids = [i for i in ids if i > EOS_IDX]
        return ' '.join(self.itos[i] for i in ids)

    def idx2sentences(self, idxs, print_special_tokens=True):
        first = idxs[0]
        if not isinstance(first, list) and (isinstance(first, (int, float)) or getattr(first, 'dim', lambda: 1)() == 0):
            return self.idx2sentence(idxs, print_special_tokens)
        return [self.idx2sentences(s, print_special_tokens) for s in idxs]
