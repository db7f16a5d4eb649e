"""Oracle for the decode side of CLaSS: beam / greedy generation from (z, c),
and the Kim-CNN classifier forward.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Beam search restates
`models/model.py:258-276,295-328,364-376` + `models/Beam.py:56-132` per sample
(SURVEY.md appendix B); greedy restates `models/model.py:277-311,348-363`.
"""
import numpy as np
import torch

from .wae import (PAD_IDX, START_IDX, EOS_IDX, DEC_H, MAX_SEQ_LEN)

NEG = -1e20     # Beam.py:69,75 sentinel


def decoder_step(p, tok, zc, h):
    """One `GRUDecoder.forward_sample` step in eval mode (decoder.py:86-109):
    x = [E[tok]; z; c]; GRU cell; logits = fc(h') (dropout is identity)."""
    H = DEC_H
    x = torch.cat([p['word_emb.weight'][tok], zc], dim=1)
    gi = x @ p['decoder.rnn.weight_ih_l0'].t() + p['decoder.rnn.bias_ih_l0']
    gh = h @ p['decoder.rnn.weight_hh_l0'].t() + p['decoder.rnn.bias_hh_l0']
    r = torch.sigmoid(gi[:, :H] + gh[:, :H])
    u = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
    n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
    h2 = (1 - u) * n + u * h
    return h2 @ p['decoder.fc.1.weight'].t() + p['decoder.fc.1.bias'], h2


def _topk_lowest_index_first(flat, k):
    """Top-k, descending, ties broken by the LOWER flat index (the rule the CUDA
    kernel implements; torch.topk's CPU tie order is unspecified)."""
    order = np.lexsort((np.arange(flat.shape[0]), -flat))
    return order[:k]


def beam_decode(p, z, c, beam_size=5, n_best=3, max_len=MAX_SEQ_LEN):
    """Returns (hyps, margins): hyps[j] is a list of n_best token-id lists (each
    starting with <start>), margins[j] the smallest score gap seen between any
    adjacent pair among the top (beam_size+1) candidates of that sample over all
    steps (used by tests to set aside genuine near-ties)."""
    V = p['decoder.fc.1.weight'].shape[0]
    B = z.shape[0]
    K = beam_size
    zc1 = torch.cat([z, c], dim=1).float()
    hyps_all, margins = [], []
    # Samples are independent: every sample runs its own K-row decoder.
    zc = zc1.repeat_interleave(K, dim=0)                    # row j*K + k
    h = zc.clone()
    tok = torch.full((B, K), PAD_IDX, dtype=torch.int64)
    tok[:, 0] = START_IDX
    score = np.zeros((B, K), dtype=np.float32)
    prev_ks = [[] for _ in range(B)]
    next_ys = [[tok[j].numpy().copy()] for j in range(B)]
    finished = [[] for _ in range(B)]
    eos_top = [False] * B
    margin = np.full(B, np.inf)

    def done(j):
        return eos_top[j] and len(finished[j]) >= n_best

    for step in range(max_len):
        logits, h = decoder_step(p, tok.reshape(-1), zc, h)
        lp = torch.log_softmax(logits.view(B, K, V), dim=2).numpy().copy()
        h = h.view(B, K, -1)
        new_tok = tok.clone()
        for j in range(B):
            if not done(j):
                w = lp[j]
                w[:, START_IDX] = NEG                           # Beam.py:69
                if prev_ks[j]:
                    cand = w + score[j][:, None]
                    for k in range(K):                          # Beam.py:73-75
                        if next_ys[j][-1][k] == EOS_IDX:
                            cand[k] = NEG
                    flat = cand.reshape(-1)
                else:
                    flat = w[0].copy()                          # Beam.py:77
                top = _topk_lowest_index_first(flat, min(K + 1, flat.shape[0]))
                vals = flat[top]
                live = vals > -1e19
                gaps = np.abs(np.diff(vals))[live[1:]]
                if gaps.size:
                    margin[j] = min(margin[j], float(gaps.min()))
                top = top[:K]
                score[j] = flat[top]
                pk = top // V
                ys = top - pk * V
                prev_ks[j].append(pk)
                next_ys[j].append(ys)
                for k in range(K):                              # Beam.py:95-98
                    if ys[k] == EOS_IDX:
                        finished[j].append((float(score[j][k]), len(next_ys[j]) - 1, k))
                if ys[0] == EOS_IDX:
                    eos_top[j] = True
                new_tok[j] = torch.from_numpy(ys)
            if prev_ks[j]:                                      # model.py:325, 387-404
                h[j] = h[j][torch.from_numpy(prev_ks[j][-1])]
        tok = new_tok
        h = h.reshape(B * K, -1)
        if all(done(j) for j in range(B)):
            break

    for j in range(B):
        fin = list(finished[j])
        i = 0
        while len(fin) < n_best:                                # Beam.py:110-117
            fin.append((float(score[j][i]), len(next_ys[j]) - 1, i))
            i += 1
        fin.sort(key=lambda a: -a[0])                           # stable
        hyps = []
        for (_, t, k) in fin[:n_best]:                          # Beam.py:124-132
            hyp = []
            for s in range(t - 1, -2, -1):
                hyp.append(int(next_ys[j][s + 1][k]))
                if s >= 0:
                    k = int(prev_ks[j][s][k])
            hyps.append(hyp[::-1])
        hyps_all.append(hyps)
        margins.append(float(margin[j]))
    return hyps_all, margins


def greedy_decode(p, z, c, max_len=MAX_SEQ_LEN):
    """`sample_G(sample_mode='greedy')`: argmax each step, PAD after <eos>, stop
    early once every row has finished.  Returns int64 [B, <=26] with the leading
    <start> column (prepend_start_idx=True)."""
    B = z.shape[0]
    zc = torch.cat([z, c], dim=1).float()
    h = zc.clone()
    tok = torch.full((B,), START_IDX, dtype=torch.int64)
    fin = torch.zeros(B, dtype=torch.bool)
    cols = [tok.clone()]
    for _ in range(max_len):
        logits, h = decoder_step(p, tok, zc, h)
        tok = torch.argmax(logits, 1)
        tok = tok.masked_fill(fin, PAD_IDX)
        fin = fin | (tok == EOS_IDX)
        cols.append(tok.clone())
        if bool(fin.all()):
            break
    return torch.stack(cols, dim=1)


def cnn_classifier_forward(p, tokens):
    """`RNN_VAE.forward_classifier` in eval mode (models/model.py:135-144,
    models/classifier.py:39-60): for widths 3,4,5 a valid conv over time of the
    embedded sequence with 100 filters of shape (w, 150), relu, max over time,
    concat (300), Linear(300 -> 2).  Dropout is identity in eval."""
    emb = p['word_emb.weight'][tokens]                         # [B, L, E]
    B, L, E = emb.shape
    feats = []
    for i, w in enumerate((3, 4, 5)):
        wt = p['classifier.conv_layers.%d.weight' % i].reshape(-1, w * E)   # [F, w*E]
        bs = p['classifier.conv_layers.%d.bias' % i]
        win = emb.unfold(1, w, 1)                               # [B, L-w+1, E, w]
        win = win.permute(0, 1, 3, 2).reshape(B, L - w + 1, w * E)
        act = torch.relu(win @ wt.t() + bs)
        feats.append(act.max(dim=1).values)
    f = torch.cat(feats, dim=1)
    return f @ p['classifier.fc.1.weight'].t() + p['classifier.fc.1.bias']
