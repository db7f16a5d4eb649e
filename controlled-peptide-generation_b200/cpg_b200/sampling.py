"""Functional wrappers over the decode / CLaSS-sampling / CNN entry points of the C ABI."""
import ctypes
import math
from ctypes import c_double, c_int, c_void_p

import numpy as np
import torch

from . import _lib
from ._lib import check, context, lib, ptr, stream_ptr

ZD, CD = 100, 2
MODE_GREEDY, MODE_CATEGORICAL = 1, 2
SOFT_MODES = {'none_softmax': 3, 'greedy_softmax': 4, 'categorical_softmax': 5}


# ------------------------------------------------------------------------------------- decode
def beam_decode(params, n_vocab, z, c, max_len=25, beam_size=5, n_best=3):
    """-> (tokens int32 [n, n_best, L+1] with -1 padding, lengths [n, n_best], scores [n, n_best])."""
    n = z.shape[0]
    dev = z.device
    toks = torch.empty(n, n_best, max_len + 1, dtype=torch.int32, device=dev)
    lens = torch.empty(n, n_best, dtype=torch.int32, device=dev)
    scores = torch.empty(n, n_best, dtype=torch.float32, device=dev)
    check(lib().cpg_beam_decode(context(dev), stream_ptr(), ptr(params), n_vocab, n, max_len,
                                ptr(z.contiguous(), torch.float32), ptr(c.contiguous(), torch.float32), beam_size,
                                n_best, ptr(toks), ptr(lens), ptr(scores)), 'cpg_beam_decode')
    return toks, lens, scores


def sample_decode(params, n_vocab, z, c, mode, max_len=25, temp=1.0, seed=0):
    """Greedy / categorical decode -> int64 [n, 1 + steps] like the reference's stacked seqIx."""
    n = z.shape[0]
    dev = z.device
    toks = torch.empty(n, max_len + 1, dtype=torch.int32, device=dev)
    steps = torch.zeros(1, dtype=torch.int32, device=dev)
    check(lib().cpg_sample_decode(context(dev), stream_ptr(), ptr(params), n_vocab, n, max_len,
                                  ptr(z.contiguous(), torch.float32), ptr(c.contiguous(), torch.float32), int(mode),
                                  float(temp), int(seed), ptr(toks), ptr(steps)), 'cpg_sample_decode')
    k = int(steps.item())
    return toks[:, :k + 1].to(torch.int64)


def soft_decode(params, n_vocab, z, c, mode, max_len=25, temp=1.0, seed=0):
    """Soft sampling (models/model.py:330-359, forward only): -> (seqIx int64 [n, 1 + steps], seqSoftIx float32
    [n, 1 + steps, V]): per step softmax(logits / temp), zeroed from the <eos> step on, fed back as a soft embedding."""
    n = z.shape[0]
    dev = z.device
    toks = torch.empty(n, max_len + 1, dtype=torch.int32, device=dev)
    soft = torch.empty(n, max_len + 1, n_vocab, dtype=torch.float32, device=dev)
    steps = torch.zeros(1, dtype=torch.int32, device=dev)
    check(lib().cpg_soft_decode(context(dev), stream_ptr(), ptr(params), n_vocab, n, max_len,
                                ptr(z.contiguous(), torch.float32), ptr(c.contiguous(), torch.float32), int(SOFT_MODES[mode]),
                                float(temp), int(seed), ptr(toks), ptr(soft), ptr(steps)), 'cpg_soft_decode')
    k = int(steps.item())
    return toks[:, :k + 1].to(torch.int64), soft[:, :k + 1].contiguous()


# ---------------------------------------------------------------------------------------- CNN
def cnn_classifier_forward(emb, conv_ws, conv_bs, fc_w, fc_b, tokens):
    B, L = tokens.shape
    V = emb.shape[0]
    dev = tokens.device
    table = torch.empty(12 * V * 100, device=dev)
    logits = torch.empty(B, 2, device=dev)
    cw = [w.contiguous() for w in conv_ws]
    check(lib().cpg_cnn_classifier_fwd(context(dev), stream_ptr(), ptr(emb.contiguous()), ptr(cw[0]),
                                       ptr(conv_bs[0].contiguous()), ptr(cw[1]), ptr(conv_bs[1].contiguous()),
                                       ptr(cw[2]), ptr(conv_bs[2].contiguous()), ptr(fc_w.contiguous()),
                                       ptr(fc_b.contiguous()), V, B, L, ptr(tokens, torch.int64), ptr(table),
                                       ptr(logits)), 'cpg_cnn_classifier_fwd')
    return logits


# -------------------------------------------------------------------------------------- CLaSS
class ClassifierSpec:
    """z-space attribute classifiers in the form the kernels take (device fp64 coefficients)."""

    def __init__(self, clfs, device):
        """clfs: ordered list of (name, coef[100] numpy, intercept, target_col)."""
        if len(clfs) > 4:
            raise ValueError('at most 4 attribute classifiers')
        self.names = [c[0] for c in clfs]
        self.target = [int(c[3]) for c in clfs]
        self.f32 = [1 if np.asarray(c[1]).dtype == np.float32 else 0 for c in clfs]
        self.coef_dev = [torch.as_tensor(np.asarray(c[1], dtype=np.float64).reshape(-1)).to(device) for c in clfs]
        n = len(clfs)
        self.n = n
        self.coef_ptrs = (c_void_p * max(n, 1))(*[t.data_ptr() for t in self.coef_dev])
        self.intercept = (c_double * max(n, 1))(*[float(np.asarray(c[2]).reshape(-1)[0]) for c in clfs])
        self.target_arr = (c_int * max(n, 1))(*self.target)
        self.f32_arr = (c_int * max(n, 1))(*self.f32)
        self.all_f32 = n > 0 and all(self.f32)

    def args(self):
        return (self.n, ctypes.cast(self.coef_ptrs, c_void_p), ctypes.cast(self.intercept, c_void_p),
                ctypes.cast(self.target_arr, c_void_p), ctypes.cast(self.f32_arr, c_void_p))


def score_accept(z, u, spec):
    """Parity mode of rejection_sample: z [n,100] fp32, u [n] fp64 (device) -> probs [n_clf, n], accum, accept."""
    n = z.shape[0]
    dev = z.device
    probs = torch.empty(max(spec.n, 1), n, dtype=torch.float64, device=dev)
    accum = torch.empty(n, dtype=torch.float64, device=dev)
    accept = torch.empty(n, dtype=torch.uint8, device=dev)
    check(lib().cpg_class_score_accept(context(dev), stream_ptr(), ptr(z.contiguous(), torch.float32),
                                       ptr(u.contiguous(), torch.float64), n, *spec.args(), ptr(probs), ptr(accum),
                                       ptr(accept)), 'cpg_class_score_accept')
    return probs, accum, accept


class GmmDevice:
    """Diagonal Gaussian mixture parameters staged on the device for sampling and log-density."""

    def __init__(self, weights, means, covs_diag, device):
        w = np.asarray(weights, dtype=np.float64)
        m = np.asarray(means, dtype=np.float64)
        cv = np.asarray(covs_diag, dtype=np.float64)
        self.K, self.D = m.shape
        if self.D != ZD:
            raise ValueError('latent dim must be 100')
        self.mean32 = torch.as_tensor(m.astype(np.float32)).to(device).contiguous()
        self.sd32 = torch.as_tensor(np.sqrt(cv).astype(np.float32)).to(device).contiguous()
        cdf = np.cumsum(w / w.sum())
        cdf[-1] = 1.0
        self.cdf32 = torch.as_tensor(cdf.astype(np.float32)).to(device).contiguous()
        prec = 1.0 / cv
        self.mean_t = torch.as_tensor(np.ascontiguousarray(m.T)).to(device)
        self.prec_t = torch.as_tensor(np.ascontiguousarray(prec.T)).to(device)
        logw = np.log(w) - 0.5 * self.D * math.log(2 * math.pi) + 0.5 * np.log(prec).sum(1)
        self.logw_norm = torch.as_tensor(logw).to(device)
        self.device = device


def class_sample(gmm, spec, n, seed, offset=0, want_z=True, want_scores=True):
    """Perf mode: n i.i.d. draws with global indices [offset, offset+n).  Returns dict with
    accept (uint8), n_accepted (device uint64 scalar) and optionally z, probs, accum, comp."""
    dev = gmm.device
    out = {'accept': torch.empty(n, dtype=torch.uint8, device=dev),
           'n_accepted': torch.zeros(1, dtype=torch.int64, device=dev)}
    out['z'] = torch.empty(n, ZD, dtype=torch.float32, device=dev) if want_z else None
    out['probs'] = torch.empty(max(spec.n, 1), n, dtype=torch.float64, device=dev) if want_scores else None
    out['accum'] = torch.empty(n, dtype=torch.float64, device=dev) if want_scores else None
    check(lib().cpg_class_sample(context(dev), stream_ptr(), ptr(gmm.mean32), ptr(gmm.sd32), ptr(gmm.cdf32), gmm.K,
                                 *spec.args(), int(seed), int(offset), int(n), ptr(out['z']), ptr(out['probs']),
                                 ptr(out['accum']), ptr(out['accept']), c_void_p(None), ptr(out['n_accepted'])),
          'cpg_class_sample')
    return out


def compact_accepted(accept, first_index=0, cap=None):
    """Ascending global indices (first_index + position) of the accepted draws -> (idx int64 [cap] device,
    count device uint64 scalar).  Only the first `count` entries of idx are meaningful."""
    n = accept.shape[0]
    dev = accept.device
    cap = n if cap is None else int(cap)
    idx = torch.empty(max(cap, 1), dtype=torch.int64, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    check(lib().cpg_compact_accepted(context(dev), stream_ptr(), ptr(accept, torch.uint8), n, int(first_index), cap, ptr(idx),
                                     ptr(count)), 'cpg_compact_accepted')
    return idx, count


def gather_rows(src, idx, index_base=0):
    m, D = idx.shape[0], src.shape[1]
    dst = torch.empty(m, D, dtype=torch.float32, device=src.device)
    if m:
        check(lib().cpg_gather_rows(context(src.device), stream_ptr(), ptr(src, torch.float32), ptr(idx, torch.int64),
                                    int(index_base), m, D, ptr(dst)), 'cpg_gather_rows')
    return dst


def class_regen(gmm, spec, seed, draw_index, want_scores=False):
    """z (and optionally scores) of the listed draws, bit-identical to class_sample(seed) at those indices."""
    m = draw_index.shape[0]
    dev = gmm.device
    z = torch.empty(m, ZD, dtype=torch.float32, device=dev)
    probs = torch.empty(max(spec.n, 1), m, dtype=torch.float64, device=dev) if want_scores else None
    accum = torch.empty(m, dtype=torch.float64, device=dev) if want_scores else None
    if m:
        n_clf, coef, icpt, tgt, f32 = spec.args()
        check(lib().cpg_class_regen(context(dev), stream_ptr(), ptr(gmm.mean32), ptr(gmm.sd32), ptr(gmm.cdf32), gmm.K, n_clf,
                                    coef, icpt, tgt, f32, int(seed), ptr(draw_index.contiguous(), torch.int64), m, ptr(z),
                                    ptr(probs), ptr(accum)), 'cpg_class_regen')
    return z, probs, accum


def gmm_logpdf(gmm, x):
    n = x.shape[0]
    out = torch.empty(n, dtype=torch.float64, device=x.device)
    check(lib().cpg_gmm_logpdf(context(x.device), stream_ptr(), ptr(x.contiguous(), torch.float32), n,
                               ptr(gmm.mean_t), ptr(gmm.prec_t), ptr(gmm.logw_norm), gmm.K, ptr(out)), 'cpg_gmm_logpdf')
    return out


def prior_logpdf(x):
    n = x.shape[0]
    out = torch.empty(n, dtype=torch.float64, device=x.device)
    check(lib().cpg_prior_logpdf(context(x.device), stream_ptr(), ptr(x.contiguous(), torch.float32), n, ptr(out)),
          'cpg_prior_logpdf')
    return out
