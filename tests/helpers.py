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
