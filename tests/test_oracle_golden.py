"""The oracle restatement (oracle/*.py) against the golden vectors produced by
the live reference (oracle/gen_golden.py -> tests/golden/).  CPU only."""
import zlib

import numpy as np
import pytest
import torch

from oracle import wae as ow
from oracle import class_sampling as oc
from oracle import decode as od
from conftest import load_golden

LOG_TO_ORACLE = dict(zip(
    ['train_z_mu_L1', 'train_z_logvar', 'train_z_logvar_L1', 'train_z_logvar_KL_penalty',
     'train_L_vae', 'train_L_vae_recon', 'train_L_vae_kl', 'train_L_wae_mmd', 'train_L_wae_mmdrf',
     'train_beta'],
    ['z_mu_l1', 'z_logvar_mean', 'logvar_l1', 'logvar_kl', 'loss', 'recon', 'kl', 'mmd', 'mmdrf',
     'beta']))


# Adam moves every weight by ~lr=1e-3 per step whatever the gradient's size, so an element whose
# gradient is rounding noise may differ by a fraction of a step: 2e-5 = 2% of one step.
PARAM_ATOL = 2e-5


def digest(name, t, n=32):
    a = torch.as_tensor(t).detach().double().reshape(-1).numpy()
    rs = np.random.RandomState(zlib.crc32(name.encode()) & 0x7fffffff)
    idx = rs.randint(0, a.size, size=min(n, a.size))
    return np.concatenate([[a.sum(), np.sqrt((a * a).sum())], a[idx]])


def params_from(npz):
    return {k: torch.from_numpy(npz[k].copy()) for k in npz.files}


def noise_from(fx, it):
    pre = 'it%d/noise/' % it
    return {k[len(pre):]: torch.from_numpy(fx[k].copy()) for k in fx.files if k.startswith(pre)}


@pytest.mark.parametrize('batch', [5, 32])
def test_wae_iterations_match_reference(batch):
    fx = load_golden('wae_b%d.npz' % batch)
    p = params_from(load_golden('params_init_v24.npz'))
    state = {}
    keys = [str(k) for k in fx['logged_keys']]
    for it in range(int(fx['n_it'])):
        tokens = torch.from_numpy(fx['it%d/tokens' % it])
        assert torch.equal(tokens, ow.synthetic_tokens(batch, 24, int(fx['token_seeds'][it])))
        noise = noise_from(fx, it)
        scal, grads, aux = ow.train_step(p, state, tokens, noise, it=it, beta=float(fx['betas'][it]))
        for j, k in enumerate(keys):
            assert scal[LOG_TO_ORACLE[k]] == pytest.approx(fx['logged'][it][j], rel=1e-4, abs=1e-7), k
        if it == 0:
            for k in ('mu', 'logvar', 'z', 'logits'):
                np.testing.assert_allclose(aux[k].detach().numpy(), fx['it0/' + k], rtol=1e-4, atol=2e-6)
        for k in ow.UNIQUE_VAE_PARAMS:
            ref = fx['it%d/param_digest/%s' % (it, k)]
            np.testing.assert_allclose(digest(k, p[k]), ref, rtol=2e-4, atol=PARAM_ATOL, err_msg=k)
    for k in ow.UNIQUE_VAE_PARAMS:
        np.testing.assert_allclose(p[k].numpy(), fx['final/param/' + k], rtol=2e-4, atol=PARAM_ATOL, err_msg=k)


def test_noise_and_tokens_regenerate_from_seeds():
    """The B=4096 golden stores seeds only; the generators must be reproducible."""
    fx = load_golden('wae_b32.npz')
    nz = ow.draw_noise(32, int(fx['noise_seeds'][0]))
    for k, v in nz.items():
        np.testing.assert_array_equal(v.numpy(), fx['it0/noise/' + k], err_msg=k)


def test_inference_forward_matches_reference():
    fx = load_golden('infer_b48.npz')
    p = params_from(load_golden('params_init_v24.npz'))
    tokens = torch.from_numpy(fx['tokens'])
    mu, lv = ow.encoder_forward(p, tokens)
    cnn = od.cnn_classifier_forward(p, tokens)
    np.testing.assert_allclose(mu.numpy(), fx['mu'], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(lv.numpy(), fx['logvar'], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(cnn.numpy(), fx['cnn_logits'], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(torch.softmax(cnn, 1).numpy(), fx['c'], rtol=1e-4, atol=1e-6)
    logits = ow.decoder_forward(p, tokens, mu, torch.softmax(cnn, 1), None)
    np.testing.assert_allclose(logits.numpy(), fx['dec_logits'], rtol=1e-4, atol=2e-6)


def test_class_rejection_sampling_bit_exact():
    fx = load_golden('class_sampling.npz')
    rs = np.random.RandomState(int(fx['draw_seed']))
    n = fx['z'].shape[0]
    z, comp = oc.gmm_sample(fx['gmm_weights'], fx['gmm_means'], fx['gmm_covs'], n, rs)
    u = rs.uniform(size=n)
    assert np.array_equal(z, fx['z'])
    assert np.array_equal(u, fx['u'])
    spec = [('amp', fx['amp_coef'], fx['amp_b'][0], 1), ('tox', fx['tox_coef'], fx['tox_b'][0], 0)]
    scores, acc = oc.rejection_accept(z, u, spec)
    assert np.array_equal(acc, fx['accepted'])
    assert fx['amp_coef'].dtype == np.float32 and fx['score_accum'].dtype == np.float32
    np.testing.assert_array_equal(scores['clfZ_amp=1'], fx['score_amp'])
    np.testing.assert_array_equal(scores['clfZ_tox=0'], fx['score_tox'])
    np.testing.assert_array_equal(scores['clfZ_prob_accum'], fx['score_accum'])


def test_log_densities_match_reference():
    fx = load_golden('class_sampling.npz')
    pts = fx['z'][:64]
    lq = oc.gmm_logpdf(pts, fx['gmm_weights'], fx['gmm_means'], fx['gmm_covs'])
    # sklearn scores a float32 point in float32 (rel ~2e-7); the oracle evaluates the same expansion in fp64
    np.testing.assert_allclose(lq, fx['logpdf_q'], rtol=2e-6)
    np.testing.assert_allclose(oc.prior_logpdf(pts), fx['logpdf_p'], rtol=1e-6)
    zz = oc.evaluate_nll_points(fx['nll_mu'], fx['nll_logvar'], fx['nll_noise'])
    assert -oc.gmm_logpdf(zz, fx['gmm_weights'], fx['gmm_means'], fx['gmm_covs']).mean() == \
        pytest.approx(float(fx['nll_q']), rel=1e-6)
    assert -oc.prior_logpdf(zz).mean() == pytest.approx(float(fx['nll_p']), rel=1e-6)


@pytest.mark.parametrize('tag,params', [('trained', 'params_trained_v24.npz'),
                                        ('init', 'params_init_v24.npz')])
def test_beam_and_greedy_match_reference(tag, params):
    fx = load_golden('decode.npz')
    p = params_from(load_golden(params))
    z, c = torch.from_numpy(fx[tag + '/z']), torch.from_numpy(fx[tag + '/c'])
    hyps, margins = od.beam_decode(p, z, c)
    ref = fx[tag + '/beam_hyps']
    for j, hs in enumerate(hyps):
        if margins[j] < 1e-5:       # genuine near-tie in the reference's own arithmetic
            continue
        for i, h in enumerate(hs):
            want = [int(t) for t in ref[j, i] if t >= 0]
            assert h == want, (j, i)
    g = od.greedy_decode(p, z, c)
    assert np.array_equal(g.numpy(), fx[tag + '/greedy'])
