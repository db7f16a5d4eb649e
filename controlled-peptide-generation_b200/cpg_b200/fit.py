"""Fitting the latent density Q(z) (diagonal Gaussian mixture, EM) and the z-space attribute classifiers (L2 logistic
regression) on the GPU -- the two scikit-learn fits of the reference's sampling set-up (density_modeling.py:64-73,
sample_pipeline.py:169-192).  The arithmetic is sklearn's (mixture/_gaussian_mixture.py, linear_model/_logistic.py
objective); what differs is documented per function."""
import ctypes
import math
import types

import numpy as np
import torch

from . import _lib
from ._lib import check, context, lib, ptr, stream_ptr

D = 100


def _logw_norm(weights, prec):
    return torch.log(weights) - 0.5 * D * math.log(2 * math.pi) + 0.5 * torch.log(prec).sum(1)


def gmm_fit_diag(x, n_components, means_init=None, weights_init=None, covs_init=None, tol=1e-3, max_iter=100,
                 reg_covar=1e-6, seed=0):
    """EM for a diagonal-covariance mixture on x (fp32 [N, 100], device).  Same loop as sklearn's
    BaseMixture.fit: E-step, M-step, stop when the mean log-likelihood changes by less than `tol`.
    Initialisation: the given means / weights / covariances, else K distinct data points, uniform weights and the
    global per-dimension variance (sklearn's default k-means start is host work and is NOT reproduced: pass its result
    as *_init to continue from it).  -> namespace(weights_, means_, covariances_, precisions_cholesky_, converged_,
    n_iter_, lower_bound_) with numpy float64 arrays, like a fitted sklearn GaussianMixture."""
    dev = x.device
    x = x.contiguous().float()
    N, K = x.shape[0], int(n_components)
    g = torch.Generator().manual_seed(seed)
    if means_init is None:
        means = x[torch.randperm(N, generator=g)[:K].to(dev)].double()
    else:
        means = torch.as_tensor(np.asarray(means_init, dtype=np.float64)).to(dev)
    weights = torch.full((K,), 1.0 / K, dtype=torch.float64, device=dev) if weights_init is None else \
        torch.as_tensor(np.asarray(weights_init, dtype=np.float64)).to(dev)
    cov = x.double().var(0, unbiased=False).clamp_min(reg_covar).repeat(K, 1) if covs_init is None else \
        torch.as_tensor(np.asarray(covs_init, dtype=np.float64)).to(dev)
    resp = torch.empty(N, K, dtype=torch.float64, device=dev)
    loglik = torch.empty(N, dtype=torch.float64, device=dev)
    w_out, m_out, c_out = torch.empty_like(weights), torch.empty_like(means), torch.empty_like(cov)
    lower_bound, converged, n_iter = -np.inf, False, 0
    for n_iter in range(1, max_iter + 1):
        prev = lower_bound
        prec = (1.0 / cov).contiguous()
        means = means.contiguous()
        logw = _logw_norm(weights, prec).contiguous()          # (named: the buffers must outlive the call)
        check(lib().cpg_gmm_em_step(context(dev), stream_ptr(), ptr(x), N, K, ptr(means), ptr(prec),
                                    ptr(logw), ctypes.c_double(reg_covar), ptr(resp), ptr(w_out),
                                    ptr(m_out), ptr(c_out), ptr(loglik)), 'cpg_gmm_em_step')
        lower_bound = float(loglik.mean())                  # the one host read per iteration (convergence test)
        weights, means, cov = w_out / w_out.sum(), m_out.clone(), c_out.clone()
        if abs(lower_bound - prev) < tol:
            converged = True
            break
    cv = cov.cpu().numpy()
    return types.SimpleNamespace(weights_=weights.cpu().numpy(), means_=means.cpu().numpy(), covariances_=cv,
                                 precisions_cholesky_=1.0 / np.sqrt(cv), converged_=converged, n_iter_=n_iter,
                                 lower_bound_=lower_bound, n_components=K, covariance_type='diag')


class DeviceLogisticRegression:
    """Binary logistic regression with sklearn's objective (C * sum_i log-loss + 1/2 |w|^2, intercept unpenalised),
    solved by damped Newton iterations whose loss / gradient / Hessian come from one pass over the data on the GPU.
    The optimum is unique: sklearn's L-BFGS stops within its tolerance of the same point.  coef_ [1, 100] and
    intercept_ [1] are float64, as sklearn's are for float64 inputs (the reference fits on mu.double())."""

    def __init__(self, C=1.0, max_iter=50, tol=1e-10):
        self.C, self.max_iter, self.tol = float(C), int(max_iter), float(tol)

    def _stats(self, x, y, w):
        n = lib().cpg_logreg_stats_len()
        out = torch.empty(n, dtype=torch.float64, device=x.device)
        wd = torch.as_tensor(w, dtype=torch.float64).to(x.device)
        check(lib().cpg_logreg_newton_stats(context(x.device), stream_ptr(), ptr(x), ptr(y), x.shape[0], ptr(wd), ptr(out)),
              'cpg_logreg_newton_stats')
        o = out.cpu().numpy()
        loss, g = o[0], o[1:1 + D + 1].copy()
        H = np.zeros((D + 1, D + 1))
        H[np.triu_indices(D + 1)] = o[1 + D + 1:]
        H = H + np.triu(H, 1).T
        pen = np.r_[np.ones(D), 0.0] / self.C                  # 1/2C |w|^2 on the coefficients only
        return loss + 0.5 * float((pen * w * w).sum()), g + pen * w, H + np.diag(pen)

    def fit(self, X, Y, device=None):
        dev = _lib.tensor_device(device)
        x = torch.as_tensor(np.asarray(X, dtype=np.float32)).to(dev).contiguous()
        y = torch.as_tensor(np.asarray(Y, dtype=np.float32)).to(dev).contiguous()
        assert x.shape[1] == D and set(np.unique(np.asarray(Y)).tolist()) <= {0.0, 1.0}
        w = np.zeros(D + 1)
        loss, g, H = self._stats(x, y, w)
        for it in range(self.max_iter):
            step = np.linalg.solve(H, g)
            t = 1.0
            while True:                                        # backtracking keeps the (convex) loss decreasing
                loss2, g2, H2 = self._stats(x, y, w - t * step)
                if loss2 <= loss or t < 1e-6:
                    break
                t *= 0.5
            w, loss, g, H = w - t * step, loss2, g2, H2
            self.n_iter_ = it + 1
            if np.abs(g).max() < self.tol * max(1.0, x.shape[0]):
                break
        self.coef_, self.intercept_, self.classes_ = w[None, :D].copy(), w[D:].copy(), np.array([0.0, 1.0])
        return self

    def decision_function(self, X):
        return np.asarray(X, dtype=np.float64) @ self.coef_[0] + self.intercept_[0]

    def predict_proba(self, X):
        p = 1.0 / (1.0 + np.exp(-self.decision_function(X)))
        return np.stack([1 - p, p], 1)

    def score(self, X, Y):
        return float(((self.decision_function(X) > 0) == (np.asarray(Y) > 0.5)).mean())
