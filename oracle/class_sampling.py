"""Oracle (numpy fp64) for the CLaSS latent sampling path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Restates
`density_modeling.py:38-80,99-108` and the scikit-learn arithmetic it delegates
to (sklearn is unpinned by amp_gen.yml:18; 1.9.0 in this image):
GaussianMixture.sample (mixture/_base.py:434-511), diag log-density
(mixture/_gaussian_mixture.py:536-542 + logsumexp) and
LogisticRegression.predict_proba (linear_model/_base.py decision_function +
expit for the binary case).
"""
import math

import numpy as np


def gmm_sample(weights, means, covs_diag, n, rng):
    """`mogQ.sample` (density_modeling.py:79-80) = sklearn GaussianMixture.sample
    for covariance_type='diag' followed by `.float()`:  component counts from ONE
    multinomial draw, then for each component in index order a block of
    `count x D` standard normals scaled by sqrt(cov) and shifted by the mean.
    Rows come out grouped by component (not shuffled).  `rng` is a numpy
    RandomState (the reference uses the global one seeded by
    sample_pipeline.py:246).  Returns (z float32 [n,D], component ids [n])."""
    counts = rng.multinomial(n, weights)
    d = means.shape[1]
    blocks, ids = [], []
    for k, cnt in enumerate(counts):
        blocks.append(means[k] + rng.standard_normal(size=(int(cnt), d)) * np.sqrt(covs_diag[k]))
        ids.append(np.full(int(cnt), k, dtype=np.int64))
    return np.vstack(blocks).astype(np.float32), np.concatenate(ids)


def gmm_sample_from_normals(means, covs_diag, comp_ids, normals):
    """Same affine map with the component ids and the standard normals given
    (what the CUDA kernel's parity mode consumes): fp64 math, then fp32."""
    x = means[comp_ids] + normals * np.sqrt(covs_diag[comp_ids])
    return x.astype(np.float32)


def lr_target_proba(z32, coef, intercept, target_col):
    """`RejSampleBase.score_clf` (density_modeling.py:43-48): binary
    LogisticRegression.predict_proba column `target_col`
    (sklearn linear_model/_base.py:366-396,429-440): s = X @ coef + b,
    p1 = expit(s), columns [1 - p1, p1].  The arithmetic type follows the dtype of
    the fitted coefficients: `build_clfZ` (sample_pipeline.py:169-186) fits on
    float32 z's, and sklearn 1.9 then keeps coef_ float32, so the whole score path
    is float32 (sgemv + float32 expit); a classifier fitted on float64 data scores
    in float64 (the float32 z is promoted)."""
    from scipy.special import expit
    coef = np.asarray(coef).reshape(-1)
    if coef.dtype == np.float32:
        s = z32.astype(np.float32) @ coef + np.float32(intercept)
    else:
        s = z32.astype(np.float64) @ coef.astype(np.float64) + float(intercept)
    p1 = expit(s)
    return p1 if target_col == 1 else 1 - p1


def rejection_accept(z32, uniforms, clfs):
    """`RejSampleBase.rejection_sample` (density_modeling.py:50-60) after the
    draw: accum = 1.0 * prod_a p_a(z) (in the scores' dtype); accepted = u < accum
    with u float64.  `clfs` is an
    ordered list of (name, coef[D], intercept, target_col).  Returns
    (scores dict as the reference names them, accepted bool[n])."""
    accum = 1.0
    scores = {}
    for name, coef, b, col in clfs:
        pr = lr_target_proba(z32, coef, b, col)
        scores['clfZ_%s=%d' % (name, col)] = pr
        accum = accum * pr
    scores['clfZ_prob_accum'] = accum
    return scores, uniforms < accum


def gmm_logpdf(x, weights, means, covs_diag):
    """`mogQ.logpdf` (density_modeling.py:75-77) = GaussianMixture.score for one
    point; batched here: log sum_k w_k N(x; mu_k, diag cov_k), fp64, using
    sklearn's expansion  sum(mu^2 prec) - 2 x.(mu prec) + x^2.prec."""
    x = np.asarray(x, dtype=np.float64)
    # a mixture fitted on float32 z's has float32 parameters and sklearn then scores in float32
    # (rel ~2e-7); the oracle evaluates the same expression with everything promoted to fp64
    weights = np.asarray(weights, dtype=np.float64)
    means = np.asarray(means, dtype=np.float64)
    covs_diag = np.asarray(covs_diag, dtype=np.float64)
    prec = 1.0 / covs_diag
    d = means.shape[1]
    quad = (means ** 2 * prec).sum(1)[None, :] - 2.0 * (x @ (means * prec).T) + (x ** 2) @ prec.T
    log_det = 0.5 * np.log(prec).sum(1)
    lp = -0.5 * (d * math.log(2 * math.pi) + quad) + log_det[None, :] + np.log(weights)[None, :]
    m = lp.max(1, keepdims=True)
    return (m + np.log(np.exp(lp - m).sum(1, keepdims=True)))[:, 0]


def prior_logpdf(z):
    """density_modeling.py:11-14, batched: -D/2 log(tau) - |z|^2 / 2."""
    z = np.asarray(z)
    d = z.shape[-1]
    return -0.5 * d * math.log(math.tau) - 0.5 * (z.astype(np.float64) ** 2).sum(-1)


def evaluate_nll_points(mu, logvar, scalar_noise):
    """The z's `evaluate_nll` scores (density_modeling.py:99-108): ONE scalar
    normal per point shared by all D dims, z = mu + exp(.5 lv) * noise (fp32)."""
    return (mu + np.exp(0.5 * logvar) * scalar_noise[:, None].astype(np.float32)).astype(np.float32)
