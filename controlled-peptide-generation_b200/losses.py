"""Loss functions with the reference's names and semantics (losses.py of
IBM/controlled-peptide-generation), evaluated by libcpg_b200 kernels.  Each differentiable loss is
a torch.autograd.Function whose forward AND backward are kernel calls (the kernels produce the
gradient together with the value), so `loss.backward()` of a reference-style training loop works
unchanged.  The fused iteration of train_vae.py does not go through this module.
"""
import math

import torch

import cfg  # access cfg.losses, like the reference
from cpg_b200 import engine
from models.mutils import PAD_IDX  # noqa: F401


class _Xent(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, sequences):
        out, dl = engine.softmax_xent(logits.detach(), sequences.contiguous(), want_grad=logits.requires_grad)
        ctx.save_for_backward(dl)
        return out[0].clone()

    @staticmethod
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        return (dl * g if dl is not None else None), None


def recon_dec(sequences, logits):
    """Mean next-token NLL over the non-<pad> targets of the batch (reference losses.py:18-31)."""
    return _Xent.apply(logits, sequences)


class _LatentStat(torch.autograd.Function):
    """index 0: kl_gaussianprior, 1: kl_gaussian_sharedmu (reference losses.py:8-15)."""

    @staticmethod
    def forward(ctx, mu, logvar, which):
        out = engine.latent_stats(mu.detach(), logvar.detach())
        ctx.save_for_backward(mu, logvar)
        ctx.which = which
        return out[which].clone()

    @staticmethod
    def backward(ctx, g):
        mu, logvar = ctx.saved_tensors
        inv_b = 1.0 / mu.shape[0]
        d_lv = 0.5 * (logvar.exp() - 1.0) * (g * inv_b)
        d_mu = mu * (g * inv_b) if ctx.which == 0 else torch.zeros_like(mu)
        return d_mu, d_lv, None


def kl_gaussianprior(mu, logvar):
    """KL(N(mu, sigma) || N(0, I)), mean over the batch."""
    return _LatentStat.apply(mu, logvar, 0)


def kl_gaussian_sharedmu(mu, logvar):
    """KL(N(mu, sigma) || N(mu, I)), mean over the batch."""
    return _LatentStat.apply(mu, logvar, 1)


def wae_mmd_gaussianprior(z, method='full_kernel'):
    """MMD between z and fresh N(0, I) samples, parametrised by cfg.losses.wae_mmd (losses.py:34-44)."""
    z_prior = torch.randn_like(z)
    cfgm = cfg.losses.wae_mmd
    if method == 'full_kernel':
        return mmd_full_kernel(z, z_prior, sigma=cfgm.sigma, kernel=cfgm.kernel)
    return mmd_rf(z, z_prior, **cfgm)


class _MmdFull(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z1, z2, sigma):
        ctx.save_for_backward(z1.detach(), z2.detach())
        ctx.sigma = sigma
        return engine.mmd_full(z1.detach(), z2.detach(), sigma)[0].clone()

    @staticmethod
    def backward(ctx, g):
        z1, z2 = ctx.saved_tensors
        return engine.mmd_full_grad(z1, z2, ctx.sigma) * g, None, None       # gradient wrt z1 only (z2 is the prior sample)


def mmd_full_kernel(z1, z2, **mmd_kwargs):
    """(sum(H) - N sum_j H_jj) / (N (N-1)), H = K11 + K22 - 2 K12: the value the reference's
    `H - torch.diag(H)` row-broadcast produces (losses.py:47-56)."""
    if mmd_kwargs.get('kernel', 'gaussian') != 'gaussian':
        raise NotImplementedError("only the 'gaussian' kernel (cfg default) is built")
    assert z1.size(0) == z2.size(0), 'expected matching sizes z1 z2'
    return _MmdFull.apply(z1, z2, float(mmd_kwargs['sigma']))


rf = {}          # cached random features, like the module global of the reference (losses.py:66)


def _get_rf(z, kernel, rf_dim, rf_resample):
    if kernel != 'gaussian':
        raise ValueError('todo implement rf for kernel ' + kernel)
    if kernel not in rf or rf_resample:
        rf_w = torch.randn((z.shape[1], rf_dim), device=z.device)
        rf_b = math.pi * 2 * torch.rand((rf_dim,), device=z.device)
        rf['gaussian'] = (rf_w, rf_b)
    rf_w, rf_b = rf['gaussian']
    assert rf_w.shape == (z.shape[1], rf_dim), 'not expecting z dim or rf_dim to change'
    return rf_w, rf_b


class _MmdRf(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z1, z2, rf_w, rf_b, sigma):
        out, dz = engine.mmd_rf(z1.detach(), z2.detach(), rf_w, rf_b, sigma, want_grad=z1.requires_grad)
        ctx.save_for_backward(dz)
        return out[0].clone()

    @staticmethod
    def backward(ctx, g):
        (dz,) = ctx.saved_tensors
        return (dz * g if dz is not None else None), None, None, None, None


def mmd_rf(z1, z2, sigma, kernel, rf_dim, rf_resample=False):
    """|mean phi(z1) - mean phi(z2)|^2 with random Fourier features (losses.py:59-63)."""
    rf_w, rf_b = _get_rf(z1, kernel, rf_dim, rf_resample)
    return _MmdRf.apply(z1, z2, rf_w, rf_b, float(sigma))


def compute_gaussian_rf(z, rf_w, rf_b, sigma, rf_dim):
    """phi(z) = cos(z W / sigma + b) sqrt(2 / R)  (feature matrix; API parity, not a hot path)."""
    return torch.cos((z @ rf_w) / sigma + rf_b) * (2.0 / rf_dim) ** 0.5


def compute_mmd_mean_rf(z, sigma, kernel, rf_dim, rf_resample=False):
    rf_w, rf_b = _get_rf(z, kernel, rf_dim, rf_resample)
    return compute_gaussian_rf(z, rf_w, rf_b, sigma, rf_dim).mean(0, keepdim=False)


def compute_mmd_kernel(x, y, sigma, kernel):
    """Dense N x M kernel matrix (API parity; the training path never materialises it)."""
    d2 = torch.cdist(x, y) ** 2
    if kernel == 'gaussian':
        return torch.exp(-d2 / sigma ** 2)
    if kernel == 'laplace':
        return torch.exp(-torch.sqrt(d2 + sigma ** 2))
    if kernel == 'energy':
        return torch.pow(d2 + sigma ** 2, -.25)
    raise ValueError('unknown kernel ' + kernel)


def zerodiag(M):
    assert M.dim() == 2 and M.size(0) == M.size(1), 'expect square matrix'
    out = M.clone()
    out.fill_diagonal_(0)
    return out
