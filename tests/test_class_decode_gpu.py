"""GPU parity of the CLaSS sampling / decode side (C ABI -> sm_100a kernels) against the oracle
and the golden fixtures of the live reference.  Token ids and accept masks bit-exact
(BASELINE.json), classifier logits 1e-4 relative."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from helpers import compare_beam
from oracle import class_sampling as oc
from oracle import decode as od
from oracle import wae as ow

pytestmark = pytest.mark.gpu
V = 24


@pytest.fixture(scope='module')
def mods():
    from cpg_b200 import engine, sampling
    return engine, sampling


def _params(name):
    fx = load_golden(name)
    return {k: torch.from_numpy(fx[k].copy()) for k in fx.files}


def _state(engine, p, dev):
    st = engine.FlatState(V, dev)
    st.load(p)
    return st


def _hyps_from(toks, lens):
    toks, lens = toks.cpu().numpy(), lens.cpu().numpy()
    return [[[int(t) for t in toks[j, i, :lens[j, i]]] for i in range(toks.shape[1])] for j in range(toks.shape[0])]


@pytest.mark.parametrize('tag,pfile', [('trained', 'params_trained_v24.npz'), ('init', 'params_init_v24.npz')])
def test_beam_and_greedy_match_reference_golden(mods, tag, pfile):
    engine, sampling = mods
    dev = torch.device('cuda')
    fx = load_golden('decode.npz')
    p = _params(pfile)
    st = _state(engine, p, dev)
    z, c = torch.from_numpy(fx[tag + '/z']).to(dev), torch.from_numpy(fx[tag + '/c']).to(dev)
    toks, lens, scores = sampling.beam_decode(st.params, V, z, c)
    got = _hyps_from(toks, lens)
    ref = [[[int(t) for t in h if t >= 0] for h in hs] for hs in fx[tag + '/beam_hyps']]
    skipped, _ = compare_beam(got, ref, fx[tag + '/beam_margin'], tag)
    assert skipped == 0                  # every golden sample is held to bit-exact token ids
    g = sampling.sample_decode(st.params, V, z, c, sampling.MODE_GREEDY)
    assert np.array_equal(g.cpu().numpy(), fx[tag + '/greedy'])


@pytest.mark.parametrize('n', [1, 6, 7, 50])
def test_beam_matches_oracle_ragged_sizes(mods, n):
    engine, sampling = mods
    dev = torch.device('cuda')
    p = _params('params_trained_v24.npz')
    st = _state(engine, p, dev)
    g = torch.Generator().manual_seed(100 + n)
    z = torch.randn(n, 100, generator=g)
    c = torch.eye(2)[torch.randint(0, 2, (n,), generator=g)]
    hyps, margins = od.beam_decode(p, z, c)
    toks, lens, scores = sampling.beam_decode(st.params, V, z.to(dev), c.to(dev))
    got = _hyps_from(toks, lens)
    compare_beam(got, hyps, margins, 'ragged n=%d' % n)
    og = od.greedy_decode(p, z, c)
    gg = sampling.sample_decode(st.params, V, z.to(dev), c.to(dev), sampling.MODE_GREEDY)
    assert np.array_equal(gg.cpu().numpy(), og.numpy())


def test_categorical_decode_is_a_valid_sampler(mods):
    """Statistical check: first-token frequencies follow softmax(logits) of the first step."""
    engine, sampling = mods
    dev = torch.device('cuda')
    p = _params('params_trained_v24.npz')
    st = _state(engine, p, dev)
    n = 20000
    z1 = torch.randn(1, 100, generator=torch.Generator().manual_seed(1))
    c1 = torch.tensor([[1.0, 0.0]])
    z, c = z1.repeat(n, 1), c1.repeat(n, 1)
    toks = sampling.sample_decode(st.params, V, z.to(dev), c.to(dev), sampling.MODE_CATEGORICAL, temp=1.0, seed=7).cpu()
    zc = torch.cat([z1, c1], 1)
    logits, _ = od.decoder_step(p, torch.tensor([ow.START_IDX]), zc, zc.clone())
    want = torch.softmax(logits[0], 0).numpy()
    freq = np.bincount(toks[:, 1].numpy(), minlength=V) / n
    assert np.abs(freq - want).max() < 4 * np.sqrt(0.25 / n) + 1e-3
    assert (toks[:, 0] == ow.START_IDX).all()
    # <pad> after <eos>
    t = toks.numpy()
    for row in t[:200]:
        e = np.where(row == ow.EOS_IDX)[0]
        if e.size:
            assert (row[e[0] + 1:] == ow.PAD_IDX).all()


def test_cnn_classifier_matches_reference_golden(mods):
    engine, sampling = mods
    dev = torch.device('cuda')
    fx = load_golden('infer_b48.npz')
    p = _params('params_init_v24.npz')
    d = lambda k: p[k].to(dev)
    logits = sampling.cnn_classifier_forward(
        d('word_emb.weight'), [d('classifier.conv_layers.%d.weight' % i) for i in range(3)],
        [d('classifier.conv_layers.%d.bias' % i) for i in range(3)], d('classifier.fc.1.weight'),
        d('classifier.fc.1.bias'), torch.from_numpy(fx['tokens']).to(dev))
    np.testing.assert_allclose(logits.cpu().numpy(), fx['cnn_logits'], rtol=1e-4, atol=1e-6)


def _spec(sampling, fx, dev, as_f64=False):
    cast = (lambda a: a.astype(np.float64)) if as_f64 else (lambda a: a)
    clfs = [('amp', cast(fx['amp_coef']), cast(fx['amp_b']), 1), ('tox', cast(fx['tox_coef']), cast(fx['tox_b']), 0)]
    return clfs, sampling.ClassifierSpec(clfs, dev)


def test_rejection_accept_matches_reference_golden(mods):
    """z, u = the reference's own draws; accept mask bit-exact, scores to float32 rounding."""
    engine, sampling = mods
    dev = torch.device('cuda')
    fx = load_golden('class_sampling.npz')
    clfs, spec = _spec(sampling, fx, dev)
    z, u = torch.from_numpy(fx['z']).to(dev), torch.from_numpy(fx['u']).to(dev)
    probs, accum, accept = sampling.score_accept(z, u, spec)
    borderline = np.abs(fx['u'] - fx['score_accum'].astype(np.float64)) < 1e-6
    got = accept.cpu().numpy().astype(bool)
    assert np.array_equal(got[~borderline], fx['accepted'][~borderline])
    assert borderline.sum() <= 2
    # fp32 dot of 100 terms in a different order than sgemv: |ds| ~ 1e-6, dp/p = (1-p) ds
    np.testing.assert_allclose(probs[0].cpu().numpy(), fx["score_amp"], rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(probs[1].cpu().numpy(), fx["score_tox"], rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(accum.cpu().numpy(), fx["score_accum"], rtol=4e-5, atol=1e-7)


def test_rejection_accept_float64_classifiers_match_oracle(mods):
    engine, sampling = mods
    dev = torch.device('cuda')
    fx = load_golden('class_sampling.npz')
    clfs, spec = _spec(sampling, fx, dev, as_f64=True)
    rs = np.random.RandomState(5)
    n = 100000
    z, _ = oc.gmm_sample(fx['gmm_weights'], fx['gmm_means'], fx['gmm_covs'], n, rs)
    u = rs.uniform(size=n)
    scores, acc = oc.rejection_accept(z, u, clfs)
    probs, accum, accept = sampling.score_accept(torch.from_numpy(z).to(dev), torch.from_numpy(u).to(dev), spec)
    assert np.array_equal(accept.cpu().numpy().astype(bool), acc)
    # 1 - expit(s) cancels for s >> 0: absolute error ~1e-16 whatever the size of the result
    np.testing.assert_allclose(accum.cpu().numpy(), scores['clfZ_prob_accum'], rtol=1e-10, atol=1e-15)


def test_log_densities_match_oracle_and_golden(mods):
    engine, sampling = mods
    dev = torch.device('cuda')
    fx = load_golden('class_sampling.npz')
    gmm = sampling.GmmDevice(fx['gmm_weights'], fx['gmm_means'], fx['gmm_covs'], dev)
    x = torch.from_numpy(fx['z'])
    lq = sampling.gmm_logpdf(gmm, x.to(dev)).cpu().numpy()
    np.testing.assert_allclose(lq, oc.gmm_logpdf(fx['z'], fx['gmm_weights'], fx['gmm_means'], fx['gmm_covs']), rtol=1e-9)
    np.testing.assert_allclose(lq[:64], fx['logpdf_q'], rtol=2e-6)
    lp = sampling.prior_logpdf(x.to(dev)).cpu().numpy()
    np.testing.assert_allclose(lp, oc.prior_logpdf(fx['z']), rtol=1e-12)
    np.testing.assert_allclose(lp[:64], fx['logpdf_p'], rtol=1e-6)


def test_perf_mode_sampler_statistics(mods):
    """Philox draws: component frequencies, per-dimension moments, acceptance rate vs the oracle on
    the same z, uniformity of the acceptance uniforms, determinism and offset-sharding."""
    engine, sampling = mods
    dev = torch.device('cuda')
    fx = load_golden('class_sampling.npz')
    clfs, spec = _spec(sampling, fx, dev)
    w, m, cv = fx['gmm_weights'], fx['gmm_means'], fx['gmm_covs']
    gmm = sampling.GmmDevice(w, m, cv, dev)
    n = 400000
    out = sampling.class_sample(gmm, spec, n, seed=11)
    z = out['z'].cpu().numpy()
    acc = out['accept'].cpu().numpy().astype(bool)
    assert int(out['n_accepted'].item()) == int(acc.sum())
    # the accept decision is consistent with the scores the kernel reports
    accum = out['accum'].cpu().numpy()
    oscores, _ = oc.rejection_accept(z, np.zeros(n), clfs)
    # float32 score path: 1 - expit(s) cancels for s >> 0, so tiny probabilities carry an absolute
    # error of a few ulp(1) = 6e-8 whatever their size
    np.testing.assert_allclose(accum, oscores['clfZ_prob_accum'], rtol=1e-4, atol=4e-7)
    assert abs(acc.mean() - accum.mean()) < 4 * np.sqrt(0.25 / n)
    # moments of the mixture
    mean_want = (w[:, None] * m).sum(0)
    var_want = (w[:, None] * (cv + m ** 2)).sum(0) - mean_want ** 2
    assert np.abs(z.mean(0) - mean_want).max() < 6 * np.sqrt(var_want.max() / n)
    np.testing.assert_allclose(z.var(0), var_want, rtol=0.03)
    # same seed -> same draws; shards by offset reproduce the un-sharded stream
    a = sampling.class_sample(gmm, spec, 1000, seed=11)
    b = sampling.class_sample(gmm, spec, 500, seed=11, offset=500)
    assert torch.equal(a['z'][500:], b['z']) and torch.equal(a['accept'][500:], b['accept'])
    assert torch.equal(a['z'], out['z'][:1000])
