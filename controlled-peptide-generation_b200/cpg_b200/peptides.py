"""Peptide-level post-processing of decoded samples on the GPU (sample_pipeline.py:210-218,312-316 of the reference):
duplicate removal and the three modlamp GlobalAnalysis descriptors H, uH, charge.

modlamp is not installed in this image (and unpinned in the reference's amp_gen.yml), so its formulas are restated
here from its published source (modlamp.descriptors: PeptideDescriptor.calculate_global / calculate_moment with
window=1000 >= len, angle=100, modality='max'; GlobalDescriptor._charge with ph=7.0, amide=True) together with its
Eisenberg consensus scale; the reference's own evaluator (evals/peptide_evals.py:18-28) carries the NORMALISED
Eisenberg scale, available here as scale='eisenberg_norm'.  The tests compare with a plain-Python restatement kept with the test infrastructure.
"""
import ctypes
from ctypes import c_void_p

import numpy as np
import torch

from . import _lib
from ._lib import check, context, lib, ptr, stream_ptr

AA = 'ACDEFGHIKLMNPQRSTVWY'
AA_INDEX = {a: i for i, a in enumerate(AA)}
SCALES = {
    # Eisenberg consensus hydrophobicity (Eisenberg et al. 1984) -- modlamp's 'eisenberg'
    'eisenberg': dict(A=0.62, C=0.29, D=-0.90, E=-0.74, F=1.19, G=0.48, H=-0.40, I=1.38, K=-1.50, L=1.06, M=0.64,
                      N=-0.78, P=0.12, Q=-0.85, R=-2.53, S=-0.18, T=-0.05, V=1.08, W=0.81, Y=0.26),
    # normalised consensus scale of the reference's PeptideEvaluator (evals/peptide_evals.py:18-25)
    'eisenberg_norm': dict(A=0.25, R=-1.80, N=-0.64, D=-0.72, C=0.04, Q=-0.69, E=-0.62, G=0.16, H=-0.40, I=0.73,
                           L=0.53, K=-1.10, M=0.26, F=0.61, P=-0.07, S=-0.26, T=-0.18, W=0.37, Y=0.02, V=0.54),
}
# pK values of modlamp.descriptors._charge
POS_PKS = {'Nterm': 9.38, 'K': 10.67, 'R': 12.10, 'H': 6.04}
NEG_PKS = {'Cterm': 2.15, 'D': 3.71, 'E': 4.15, 'C': 8.14, 'Y': 10.10}
NEG_PKS_AMIDE = dict(NEG_PKS, Cterm=15.0)


def charge_tables(ph=7.0, amide=True):
    """-> (side-chain partial charge per residue code [20] float64, N-terminus + C-terminus charge)."""
    neg = NEG_PKS_AMIDE if amide else NEG_PKS
    q = np.zeros(20, dtype=np.float64)
    ends = 0.0
    for aa, pk in POS_PKS.items():
        v = 10.0 ** pk / (10.0 ** pk + 10.0 ** ph)
        if aa == 'Nterm':
            ends += v
        else:
            q[AA_INDEX[aa]] += v
    for aa, pk in neg.items():
        v = 10.0 ** ph / (10.0 ** pk + 10.0 ** ph)
        if aa == 'Cterm':
            ends -= v
        else:
            q[AA_INDEX[aa]] -= v
    return q, ends


def token_to_residue(idx2word, n_tokens):
    """int8 [n_tokens]: residue code of every vocabulary entry, -1 for specials / unknown symbols."""
    out = np.full(n_tokens, -1, dtype=np.int8)
    for t in range(n_tokens):
        out[t] = AA_INDEX.get(str(idx2word(t)).strip().upper(), -1)
    return out


def vocabulary_words(dataset, n_tokens):
    """The word of every token id, asked of the dataset itself one token at a time (works with the reference's dataset
    class and with any shim that offers idx2sentences)."""
    return [str(dataset.idx2sentences([[t]], print_special_tokens=True)[0]) for t in range(n_tokens)]


def rows_to_sentences(tokens, lens, words, n_special=4, fallback=None):
    """dataset.idx2sentences([row[:n] for row, n in zip(tokens, lens)], print_special_tokens=False) for a host matrix of
    decoded rows (data_processing/dataset.py:288-300 of the reference: special ids dropped, words joined by one space)
    without a Python loop over tokens: byte matrix [n, 2 W] with the kept characters compacted to the even positions,
    viewed as fixed-width strings.  Needs every non-special word to be one ASCII character (amino-acid vocabularies);
    otherwise `fallback(rows)` is called with the list-of-lists form."""
    tok = np.asarray(tokens)
    ln = np.asarray(lens).astype(np.int64)
    n, W = tok.shape
    single = all(len(w) == 1 and ord(w) < 128 for w in words[n_special:])
    if not single or n == 0:
        rows = [r[:k] for r, k in zip(tok.tolist(), ln.tolist())]
        if fallback is None:
            return [' '.join(words[i] for i in r if i >= n_special) for r in rows]
        return fallback(rows)
    assert len(words) <= 256 and W < 128
    lut = np.zeros(256, dtype=np.uint8)
    for i, w in enumerate(words):
        lut[i] = ord(w) if i >= n_special else 0
    t = np.clip(tok, 0, 255).astype(np.uint8)
    ar = np.arange(W, dtype=np.uint8)[None, :]
    keep = (t >= n_special) & (t < len(words)) & (ar < np.clip(ln, 0, W).astype(np.uint8)[:, None])
    cnt = keep.sum(axis=1, dtype=np.uint8)[:, None]
    out = np.empty((n, 2 * W), dtype=np.uint8)
    if bool((keep == (ar < cnt)).all()):
        # no special id inside the kept part of any row (what a decoder emits): positions stay where they are
        ch = lut[t]
        ch *= keep
        out[:, 0::2] = ch
        out[:, 1::2] = (ar + 1 < cnt).view(np.uint8) << 5          # ' ' after every kept token but the last
    else:
        out[:] = 0
        pos = np.cumsum(keep, axis=1, dtype=np.int64) - 1          # output slot of every kept token
        r, c = np.nonzero(keep)
        p = pos[r, c]
        out[r, 2 * p] = lut[t[r, c]]
        out[r, 2 * p + 1] = 32
        has = cnt[:, 0] > 0
        out[np.nonzero(has)[0], 2 * cnt[has, 0].astype(np.int64) - 1] = 0
    return [b.decode('ascii') for b in out.view('S%d' % (2 * W)).ravel().tolist()]


def descriptors_from_tokens(tokens, aa_of_token, scale='eisenberg', ph=7.0, amide=True, angle=100.0):
    """tokens: int32 [n, W] on the device (decoded rows, -1 padded).  -> (H, uH, charge, length) device tensors."""
    n, W = tokens.shape
    dev = tokens.device
    hyd = np.array([SCALES[scale][a] for a in AA], dtype=np.float32)
    q, ends = charge_tables(ph, amide)
    a2t = np.ascontiguousarray(aa_of_token, dtype=np.int8)
    H = torch.empty(n, device=dev)
    uH = torch.empty(n, device=dev)
    ch = torch.empty(n, device=dev)
    ln = torch.empty(n, dtype=torch.int32, device=dev)
    if n == 0:
        return H, uH, ch, ln
    check(lib().cpg_peptide_descriptors(context(dev), stream_ptr(), ptr(tokens.contiguous(), torch.int32), n, W,
                                        a2t.ctypes.data_as(c_void_p), len(a2t), hyd.ctypes.data_as(c_void_p),
                                        q.ctypes.data_as(c_void_p), ctypes.c_double(ends), float(angle), ptr(H), ptr(uH),
                                        ptr(ch), ptr(ln)), 'cpg_peptide_descriptors')
    return H, uH, ch, ln


def descriptors_from_strings(seqs, device=None, **kw):
    """List of amino-acid strings ('' allowed: NaN H / uH) -> numpy H, uH, charge (what compute_modlamp stores)."""
    dev = _lib.tensor_device(device)
    n = len(seqs)
    if n == 0:
        z = np.zeros(0, dtype=np.float32)
        return z, z.copy(), z.copy()
    W = max(1, max(len(s) for s in seqs))
    tok = np.full((n, W), -1, dtype=np.int32)
    for i, s in enumerate(seqs):
        for j, ch in enumerate(s):
            tok[i, j] = AA_INDEX.get(ch.upper(), 63)              # unknown letters map to a non-residue code
    H, uH, ch, _ = descriptors_from_tokens(torch.from_numpy(tok).to(dev), np.arange(20, dtype=np.int8), **kw)
    return H.cpu().numpy(), uH.cpu().numpy(), ch.cpu().numpy()


def dedup_rows(tokens):
    """tokens int32 [n, W] (device) -> (first_index int32 [n], is_first uint8 [n]): pandas drop_duplicates()
    keeps exactly the rows with is_first == 1."""
    n, W = tokens.shape
    dev = tokens.device
    first = torch.empty(n, dtype=torch.int32, device=dev)
    flag = torch.empty(n, dtype=torch.uint8, device=dev)
    if n:
        check(lib().cpg_dedup_rows(context(dev), stream_ptr(), ptr(tokens.contiguous(), torch.int32), n, W, ptr(first),
                                   ptr(flag)), 'cpg_dedup_rows')
    return first, flag
