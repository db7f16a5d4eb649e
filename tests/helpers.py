"""Shared helpers for the parity tests (CUDA path vs oracle / golden fixtures)."""
import zlib

import numpy as np
import torch

from oracle import wae as ow

# Adam moves every weight by about lr=1e-3 per step whatever the gradient's size; when |g| is
# rounding noise (<~1e-7, comparable to Adam's eps=1e-8 after bias correction) the update is
# decided by that noise.  Post-step parameters are therefore compared with an absolute
# tolerance of 2% of a step, and elements whose reference gradient is noise are set aside.
PARAM_ATOL = 2e-5
GRAD_NOISE = 1e-6


def dev_noise(noise, device):
    return {k: v.to(device).contiguous() for k, v in noise.items()}


def digest(name, t, n=32):
    a = torch.as_tensor(t).detach().double().reshape(-1).cpu().numpy()
    rs = np.random.RandomState(zlib.crc32(name.encode()) & 0x7fffffff)
    idx = rs.randint(0, a.size, size=min(n, a.size))
    return np.concatenate([[a.sum(), np.sqrt((a * a).sum())], a[idx]])


def rel_err(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def clipped(ograds, grad_norm, max_norm=5.0):
    """Oracle raw grads -> what clip_grad_norm_ leaves in .grad (embedding scaled twice)."""
    coef = min(1.0, max_norm / (grad_norm + 1e-6))
    return {k: g * (coef * coef if k == 'word_emb.weight' else coef) for k, g in ograds.items()}


def assert_params_close(got, want, grads_ref, what='', max_outliers=0, outlier_cap=2e-4):
    """`max_outliers` elements per tensor may exceed the tolerance by up to `outlier_cap`
    (a fifth of one Adam step) -- used when several iterations compound the eps-regime noise."""
    for k in ow.UNIQUE_VAE_PARAMS:
        g = torch.as_tensor(got[k]).detach().cpu()
        w = torch.as_tensor(want[k]).detach().cpu()
        d = (g - w).abs()
        bad = d > (PARAM_ATOL + 2e-4 * w.abs())
        if grads_ref is not None:
            bad &= torch.as_tensor(grads_ref[k]).detach().cpu().abs() > GRAD_NOISE
        nbad = int(bad.sum())
        worst = float(d[bad].max()) if nbad else 0.0
        assert nbad <= max_outliers and worst <= max(outlier_cap, PARAM_ATOL), \
            '%s %s: %d elements differ, max %.3g' % (what, k, nbad, worst)


# Beam search: a sample is set aside only when the reference's own fp32 arithmetic saw a score gap below this
# between a kept and a dropped candidate (the oracle-vs-reference test uses the same bar; the smallest gap in the
# committed goldens is 1.3e-5, so none of them is set aside).
BEAM_TIE_MARGIN = 1e-5


def compare_beam(got, ref, margins, what='', n_hyps=None):
    """got[j][i] / ref[j][i]: token-id lists.  Every sample with margin >= BEAM_TIE_MARGIN must agree on every
    hypothesis compared; near-ties are counted, and how many of them actually differ is reported."""
    skipped = skipped_differ = 0
    for j, hs in enumerate(got):
        k = len(hs) if n_hyps is None else n_hyps
        same = all(list(hs[i]) == list(ref[j][i]) for i in range(k))
        if margins[j] < BEAM_TIE_MARGIN:
            skipped += 1
            skipped_differ += 0 if same else 1
            continue
        assert same, '%s: sample %d (margin %.3g) differs: %s vs %s' % (what, j, margins[j], hs[:k], ref[j][:k])
    print('beam %s: %d samples, %d near-ties set aside (margin < %g), %d of those differ'
          % (what, len(got), skipped, BEAM_TIE_MARGIN, skipped_differ))
    return skipped, skipped_differ
