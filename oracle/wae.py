"""Oracle (torch-CPU fp32) for one WAE phase-1 training iteration.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Restates, with explicit gate
equations and with every random tensor passed in as an argument, what the
reference computes in `train_vae.py:24-42`.  Gradients come from torch autograd
over this explicit forward, so the backward oracle is the derivative of the
restated forward, not a second hand-written formula.

Parameter containers are plain dicts keyed by the reference's state_dict names
(SURVEY.md section 5), e.g. ``p['encoder.rnn.weight_hh_l0_reverse']``.
"""
import math
from collections import OrderedDict

import numpy as np
import torch

# models/mutils.py:5-8
UNK_IDX, PAD_IDX, START_IDX, EOS_IDX = 0, 1, 2, 3

# cfg.py:258-281 (model geometry); losses cfg.py:247-255
EMB_DIM, ENC_H, Z_DIM, C_DIM = 150, 80, 100, 2
DEC_H = Z_DIM + C_DIM
MAX_SEQ_LEN = 25
RF_DIM, MMD_SIGMA = 500, 7.0

# Order in which `RNN_VAE.vae_params()` yields tensors (models/model.py:88-94):
# word_emb, encoder.*, decoder.* -- and decoder.emb IS word_emb (model.py:63), so
# the embedding matrix is yielded twice.
VAE_PARAM_ORDER = (
    'word_emb.weight',
    'encoder.rnn.weight_ih_l0', 'encoder.rnn.weight_hh_l0',
    'encoder.rnn.bias_ih_l0', 'encoder.rnn.bias_hh_l0',
    'encoder.rnn.weight_ih_l0_reverse', 'encoder.rnn.weight_hh_l0_reverse',
    'encoder.rnn.bias_ih_l0_reverse', 'encoder.rnn.bias_hh_l0_reverse',
    'encoder.q_mu.weight', 'encoder.q_mu.bias',
    'encoder.q_logvar.weight', 'encoder.q_logvar.bias',
    'word_emb.weight',  # decoder.emb.weight alias -> duplicate entry
    'decoder.rnn.weight_ih_l0', 'decoder.rnn.weight_hh_l0',
    'decoder.rnn.bias_ih_l0', 'decoder.rnn.bias_hh_l0',
    'decoder.fc.1.weight', 'decoder.fc.1.bias',
)
UNIQUE_VAE_PARAMS = tuple(OrderedDict.fromkeys(VAE_PARAM_ORDER))


def param_shapes(n_vocab):
    """Shapes of the VAE parameters for vocabulary size ``n_vocab``."""
    g_e, g_d = 3 * ENC_H, 3 * DEC_H
    return OrderedDict([
        ('word_emb.weight', (n_vocab, EMB_DIM)),
        ('encoder.rnn.weight_ih_l0', (g_e, EMB_DIM)),
        ('encoder.rnn.weight_hh_l0', (g_e, ENC_H)),
        ('encoder.rnn.bias_ih_l0', (g_e,)),
        ('encoder.rnn.bias_hh_l0', (g_e,)),
        ('encoder.rnn.weight_ih_l0_reverse', (g_e, EMB_DIM)),
        ('encoder.rnn.weight_hh_l0_reverse', (g_e, ENC_H)),
        ('encoder.rnn.bias_ih_l0_reverse', (g_e,)),
        ('encoder.rnn.bias_hh_l0_reverse', (g_e,)),
        ('encoder.q_mu.weight', (Z_DIM, 2 * ENC_H)),
        ('encoder.q_mu.bias', (Z_DIM,)),
        ('encoder.q_logvar.weight', (Z_DIM, 2 * ENC_H)),
        ('encoder.q_logvar.bias', (Z_DIM,)),
        ('decoder.rnn.weight_ih_l0', (g_d, EMB_DIM + DEC_H)),
        ('decoder.rnn.weight_hh_l0', (g_d, DEC_H)),
        ('decoder.rnn.bias_ih_l0', (g_d,)),
        ('decoder.rnn.bias_hh_l0', (g_d,)),
        ('decoder.fc.1.weight', (n_vocab, DEC_H)),
        ('decoder.fc.1.bias', (n_vocab,)),
    ])


def random_params(n_vocab, seed=0, scale=1.0):
    """Synthetic parameters with the init ranges torch uses (uniform +-1/sqrt(H)
    for GRU/Linear, N(0,1) embedding with a zero PAD row, models/model.py:47)."""
    g = torch.Generator().manual_seed(seed)
    p = OrderedDict()
    for name, shape in param_shapes(n_vocab).items():
        if name == 'word_emb.weight':
            w = torch.randn(shape, generator=g)
            w[PAD_IDX] = 0.0
        else:
            if name.startswith('encoder.rnn'):
                k = 1.0 / math.sqrt(ENC_H)
            elif name.startswith('decoder.rnn'):
                k = 1.0 / math.sqrt(DEC_H)
            elif name.startswith('encoder.q_'):
                k = 1.0 / math.sqrt(2 * ENC_H)
            else:
                k = 1.0 / math.sqrt(DEC_H)
            w = (torch.rand(shape, generator=g) * 2 - 1) * k * scale
        p[name] = w.float()
    return p


def synthetic_tokens(batch, n_vocab, seed=1238, max_len=MAX_SEQ_LEN):
    """rows ``[<start>, aa * len, <eos>, <pad>...]`` with
    len ~ U{5..23}, amino-acid ids U{4..V-1}; int64 [B, 25]."""
    g = torch.Generator().manual_seed(seed)
    toks = torch.full((batch, max_len), PAD_IDX, dtype=torch.int64)
    lens = torch.randint(5, max_len - 1, (batch,), generator=g)
    body = torch.randint(4, n_vocab, (batch, max_len), generator=g)
    for b in range(batch):
        n = int(lens[b])
        toks[b, 0] = START_IDX
        toks[b, 1:1 + n] = body[b, :n]
        toks[b, 1 + n] = EOS_IDX
    return toks


# --------------------------------------------------------------------------- GRU
def gru_sequence(gi, h0, w_hh, b_hh, reverse=False):
    """torch.nn.GRU recurrence with gate order (r, z, n), given the input-side
    pre-activations ``gi = x @ W_ih^T + b_ih`` of shape [B, L, 3H].

    r = sig(gi_r + W_hr h + b_hr); z = sig(gi_z + W_hz h + b_hz)
    n = tanh(gi_n + r * (W_hn h + b_hn)); h' = (1 - z) * n + z * h
    (models/encoder.py:25-30,42 and models/decoder.py:40,77 call nn.GRU.)
    Returns all hidden states [B, L, H] in input time order.
    """
    B, L, G = gi.shape
    H = G // 3
    h = h0
    outs = [None] * L
    order = range(L - 1, -1, -1) if reverse else range(L)
    for t in order:
        gh = h @ w_hh.t() + b_hh
        r = torch.sigmoid(gi[:, t, :H] + gh[:, :H])
        zg = torch.sigmoid(gi[:, t, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[:, t, 2 * H:] + r * gh[:, 2 * H:])
        h = (1.0 - zg) * n + zg * h
        outs[t] = h
    return torch.stack(outs, dim=1)


def encoder_forward(p, tokens):
    """models/model.py:96-105 + models/encoder.py:38-52.  All 25 positions are
    fed, PAD included (no packing); bi-GRU final states concatenated (fwd final =
    after t=24, bwd final = after t=0), then q_mu / q_logvar."""
    emb = p['word_emb.weight'][tokens]                       # [B, L, 150]
    B = tokens.shape[0]
    h0 = torch.zeros(B, ENC_H, dtype=emb.dtype)
    gi_f = emb @ p['encoder.rnn.weight_ih_l0'].t() + p['encoder.rnn.bias_ih_l0']
    gi_b = emb @ p['encoder.rnn.weight_ih_l0_reverse'].t() + p['encoder.rnn.bias_ih_l0_reverse']
    hs_f = gru_sequence(gi_f, h0, p['encoder.rnn.weight_hh_l0'], p['encoder.rnn.bias_hh_l0'])
    hs_b = gru_sequence(gi_b, h0, p['encoder.rnn.weight_hh_l0_reverse'],
                        p['encoder.rnn.bias_hh_l0_reverse'], reverse=True)
    h = torch.cat([hs_f[:, -1], hs_b[:, 0]], dim=1)           # [B, 160]
    mu = h @ p['encoder.q_mu.weight'].t() + p['encoder.q_mu.bias']
    logvar = h @ p['encoder.q_logvar.weight'].t() + p['encoder.q_logvar.bias']
    return mu, logvar


def reparameterize(mu, logvar, eps):
    """models/model.py:107-112."""
    return mu + torch.exp(logvar / 2) * eps


def word_dropout(tokens, drop_mask):
    """models/decoder.py:117-133: positions with mask==1 become <unk> (every
    position is eligible, <start> and <pad> included)."""
    out = tokens.clone()
    out[drop_mask.bool()] = UNK_IDX
    return out


def decoder_hidden_states(p, dec_tokens, z, c):
    """GRU part of models/decoder.py:56-77: x_t = [E[w_t]; z; c], h0 = [z; c]."""
    emb = p['word_emb.weight'][dec_tokens]                    # [B, L, 150]
    zc = torch.cat([z, c], dim=1)                             # [B, 102]
    x = torch.cat([emb, zc.unsqueeze(1).expand(-1, emb.shape[1], -1)], dim=2)
    gi = x @ p['decoder.rnn.weight_ih_l0'].t() + p['decoder.rnn.bias_ih_l0']
    return gru_sequence(gi, zc, p['decoder.rnn.weight_hh_l0'], p['decoder.rnn.bias_hh_l0'])


def decoder_forward(p, dec_tokens, z, c, out_keep=None, p_out_dropout=0.3):
    """models/decoder.py:56-84.  ``out_keep`` is the Bernoulli(1-p) keep mask of
    nn.Dropout (train mode: kept activations scaled by 1/(1-p)); None = eval."""
    hs = decoder_hidden_states(p, dec_tokens, z, c)           # [B, L, 102]
    if out_keep is not None:
        hs = hs * out_keep.to(hs.dtype) * (1.0 / (1.0 - p_out_dropout))
    return hs @ p['decoder.fc.1.weight'].t() + p['decoder.fc.1.bias']


# ------------------------------------------------------------------------ losses
def recon_dec(tokens, logits):
    """losses.py:18-31: targets = tokens shifted left with PAD appended; mean NLL
    over the non-PAD targets of the whole batch."""
    B, L, V = logits.shape
    tgt = torch.cat([tokens[:, 1:], torch.full((B, 1), PAD_IDX, dtype=tokens.dtype)], dim=1)
    lsm = logits - torch.logsumexp(logits, dim=2, keepdim=True)
    nll = -lsm.gather(2, tgt.unsqueeze(2)).squeeze(2)
    keep = tgt != PAD_IDX
    return (nll * keep).sum() / keep.sum()


def kl_gaussianprior(mu, logvar):
    """losses.py:8-10."""
    return torch.mean(0.5 * torch.sum(logvar.exp() + mu ** 2 - 1 - logvar, 1))


def kl_gaussian_sharedmu(mu, logvar):
    """losses.py:13-15."""
    return torch.mean(0.5 * torch.sum(logvar.exp() - 1 - logvar, 1))


def _gauss_gram_sums(x, y, sigma, chunk=256):
    """Sum and diagonal of K = exp(-|x_i - y_j|^2 / sigma^2) (losses.py:96-108),
    evaluated by row chunks with the reference's (x_i - y_j)**2 arithmetic so the
    [N, N, D] broadcast is never materialised."""
    total = torch.zeros((), dtype=torch.float64)
    diag = []
    for s in range(0, x.shape[0], chunk):
        d2 = ((x[s:s + chunk, None, :] - y[None, :, :]) ** 2).sum(2)
        k = torch.exp(-d2 / sigma ** 2)
        total = total + k.double().sum()
        idx = torch.arange(s, min(s + chunk, x.shape[0]))
        diag.append(k[idx - s, idx])
    return total, torch.cat(diag)


def mmd_full_kernel(z1, z2, sigma=MMD_SIGMA):
    """losses.py:47-56 *as executed*: ``H - torch.diag(H)`` subtracts the diagonal
    VECTOR broadcast over rows, so the value is
    (sum(H) - N * sum_j H_jj) / (N (N-1)),  H = K11 + K22 - 2 K12.
    Accumulated in fp64 (the reference's fp32 pairwise sum agrees to ~1e-6 rel;
    pinned against the live reference in tests/golden)."""
    n = z1.shape[0]
    s11, d11 = _gauss_gram_sums(z1, z1, sigma)
    s22, d22 = _gauss_gram_sums(z2, z2, sigma)
    s12, d12 = _gauss_gram_sums(z1, z2, sigma)
    h_sum = s11 + s22 - 2.0 * s12
    h_diag = (d11 + d22 - 2.0 * d12).double().sum()
    return ((h_sum - n * h_diag) / (n * (n - 1))).float()


def gaussian_rf(z, rf_w, rf_b, sigma=MMD_SIGMA, rf_dim=RF_DIM):
    """losses.py:90-93."""
    return torch.cos((z @ rf_w) / sigma + rf_b) * (2.0 / rf_dim) ** 0.5


def mmd_rf(z1, z2, rf_w, rf_b, sigma=MMD_SIGMA, rf_dim=RF_DIM):
    """losses.py:59-63,69-87."""
    mu1 = gaussian_rf(z1, rf_w, rf_b, sigma, rf_dim).mean(0)
    mu2 = gaussian_rf(z2, rf_w, rf_b, sigma, rf_dim).mean(0)
    return ((mu1 - mu2) ** 2).sum()


def anneal_beta(it, start_val=1.0, end_val=2.0, start_iter=0, end_iter=40000):
    """utils.py:51-61 with cfg.vae.beta (cfg.py:177-188)."""
    if it < start_iter:
        return start_val
    if it >= end_iter:
        return end_val
    return start_val + (end_val - start_val) * (it - start_iter) / (end_iter - start_iter)


# ------------------------------------------------------------- one full iteration
def wae_forward_losses(p, tokens, noise, beta=1.0, lambda_l1=0.0, lambda_kl=1e-3,
                       z_regu='mmdrf', sigma=MMD_SIGMA, rf_dim=RF_DIM, p_out_dropout=0.3,
                       with_full_mmd=True):
    """train_vae.py:24-37.  ``noise`` carries every random tensor of the
    iteration: eps [B,100], c [B,2] one-hot,
    word_drop [B,25] {0,1}, out_keep [B,25,102] {0,1}, z_prior_full [B,100],
    z_prior_rf [B,100], rf_w [100,R], rf_b [R]."""
    mu, logvar = encoder_forward(p, tokens)
    z = reparameterize(mu, logvar, noise['eps'])
    c = noise['c']
    dec_in = word_dropout(tokens, noise['word_drop'])
    logits = decoder_forward(p, dec_in, z, c, noise['out_keep'], p_out_dropout)
    out = OrderedDict()
    out['recon'] = recon_dec(tokens, logits)
    out['kl'] = kl_gaussianprior(mu, logvar)
    if with_full_mmd:
        if z_regu == 'mmd':                    # train_vae.py:29-31: the full-kernel MMD is in the loss and differentiated
            out['mmd'] = mmd_full_kernel(z, noise['z_prior_full'], sigma)
        else:
            with torch.no_grad():
                out['mmd'] = mmd_full_kernel(z.detach(), noise['z_prior_full'], sigma)
    out['mmdrf'] = mmd_rf(z, noise['z_prior_rf'], noise['rf_w'], noise['rf_b'], sigma, rf_dim)
    out['logvar_l1'] = logvar.abs().sum(1).mean(0)
    out['logvar_kl'] = kl_gaussian_sharedmu(mu, logvar)
    regu = {'kl': out['kl'], 'mmdrf': out['mmdrf']}
    if with_full_mmd:
        regu['mmd'] = out['mmd']
    out['loss'] = (out['recon'] + beta * regu[z_regu] + lambda_l1 * out['logvar_l1']
                   + lambda_kl * out['logvar_kl'])
    out['z_mu_l1'] = mu.detach().abs().mean()
    out['z_logvar_mean'] = logvar.detach().mean()
    aux = {'mu': mu, 'logvar': logvar, 'z': z, 'logits': logits}
    return out, aux


def clip_and_adam(p, grads, state, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, max_norm=5.0):
    """train_vae.py:15,41-42 with the duplicated embedding of `vae_params()`:
    (i) its squared grad norm enters the total twice, (ii) `clip_grad_norm_`
    scales its grad once per list entry (coef**2 overall), (iii) Adam performs two
    sequential updates on it per iteration with shared state (step += 2).
    ``state`` maps name -> dict(step, m, v) and is updated in place; ``p`` and
    ``grads`` are updated in place.  Returns the pre-clip total norm."""
    norms = [torch.linalg.vector_norm(grads[name]) for name in VAE_PARAM_ORDER]
    total = torch.linalg.vector_norm(torch.stack(norms))
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for name in VAE_PARAM_ORDER:            # duplicates scaled twice, as in torch
        grads[name].mul_(coef)
    b1, b2 = betas
    for name in VAE_PARAM_ORDER:            # duplicates stepped twice
        st = state.setdefault(name, {'step': 0, 'm': torch.zeros_like(p[name]),
                                     'v': torch.zeros_like(p[name])})
        g = grads[name]
        st['step'] += 1
        st['m'].lerp_(g, 1 - b1)
        st['v'].mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1 = 1 - b1 ** st['step']
        bc2 = 1 - b2 ** st['step']
        step_size = lr / bc1
        denom = (st['v'].sqrt() / math.sqrt(bc2)).add_(eps)
        p[name].addcdiv_(st['m'], denom, value=-step_size)
    return total


def train_step(p, state, tokens, noise, it=0, lr=1e-3, max_norm=5.0, beta=None, **loss_kw):
    """One iteration of train_vae.py:24-42 on explicit noise.  Mutates p/state.
    Returns (losses dict of python floats, grads dict, grad_norm)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()
              if k in UNIQUE_VAE_PARAMS}
    if beta is None:
        beta = anneal_beta(it)
    out, aux = wae_forward_losses(leaves, tokens, noise, beta=beta, **loss_kw)
    out['loss'].backward()
    grads = {k: (leaves[k].grad if leaves[k].grad is not None else torch.zeros_like(leaves[k]))
             for k in UNIQUE_VAE_PARAMS}
    grads['word_emb.weight'][PAD_IDX] = 0.0        # nn.Embedding(padding_idx=1), model.py:47
    raw = {k: g.clone() for k, g in grads.items()}
    with torch.no_grad():
        gn = clip_and_adam(p, grads, state, lr=lr, max_norm=max_norm)
    scal = {k: float(v.detach()) for k, v in out.items()}
    scal['beta'] = float(beta)
    scal['grad_norm'] = float(gn)
    return scal, raw, aux


def draw_noise(batch, seed, n_rf=RF_DIM, with_rf=True):
    """Synthetic noise bundle with the reference's distributions (appendix A)."""
    g = torch.Generator().manual_seed(seed)
    rs = np.random.RandomState(seed)
    noise = {
        'eps': torch.randn(batch, Z_DIM, generator=g),
        'c': torch.from_numpy(rs.multinomial(1, [0.5, 0.5], batch).astype('float32')),
        'word_drop': torch.from_numpy(rs.binomial(1, 0.3, (batch, MAX_SEQ_LEN)).astype('uint8')),
        'out_keep': (torch.rand(batch, MAX_SEQ_LEN, DEC_H, generator=g) >= 0.3).to(torch.uint8),
        'z_prior_full': torch.randn(batch, Z_DIM, generator=g),
        'z_prior_rf': torch.randn(batch, Z_DIM, generator=g),
    }
    if with_rf:
        g2 = torch.Generator().manual_seed(977)
        noise['rf_w'] = torch.randn(Z_DIM, n_rf, generator=g2)
        noise['rf_u'] = torch.rand(n_rf, generator=g2)
        noise['rf_b'] = math.pi * 2 * noise['rf_u']          # losses.py:76
    return noise
