"""GPU parity of the post-accept stages of a CLaSS sampling round (compaction, re-generation, gather, dedup,
descriptors) against their CPU restatements, and the reference-shaped entry point sample_pipeline.main driven end to
end on synthetic states files."""
import os
import types

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import peptides as op

pytestmark = pytest.mark.gpu
V = 24
AA = 'ACDEFGHIKLMNPQRSTVWY'
WORDS = ['<unk>', '<pad>', '<start>', '<eos>'] + list(AA)


class Dataset:
    """Stand-in for AttributeDataLoader (data_processing/dataset.py:285-300): token ids 4..23 are the amino acids."""
    def next_batch(self, name):
        raise AssertionError('not used')

    def idx2sentence(self, idxs, print_special_tokens=True):
        return self.idx2sentences([idxs], print_special_tokens)[0]

    def idx2sentences(self, seqs, print_special_tokens=True):
        return [' '.join(WORDS[int(i)] for i in s if print_special_tokens or int(i) > 3) for s in seqs]


@pytest.fixture(scope='module')
def mods():
    from cpg_b200 import peptides, sampling
    return sampling, peptides


@pytest.mark.parametrize('n', [1, 31, 2048, 2049, 100003])
def test_compaction_gather_and_regeneration(mods, n):
    sampling, _ = mods
    dev = torch.device('cuda')
    fx = load_golden('class_sampling.npz')
    clfs = [('amp', fx['amp_coef'], fx['amp_b'], 1), ('tox', fx['tox_coef'], fx['tox_b'], 0)]
    spec = sampling.ClassifierSpec(clfs, dev)
    gmm = sampling.GmmDevice(fx['gmm_weights'], fx['gmm_means'], fx['gmm_covs'], dev)
    off = 12345
    full = sampling.class_sample(gmm, spec, n, seed=5, offset=off)
    flags = sampling.class_sample(gmm, spec, n, seed=5, offset=off, want_z=False, want_scores=False)
    assert torch.equal(full['accept'], flags['accept'])
    acc = full['accept'].cpu().numpy().astype(bool)
    idx, cnt = sampling.compact_accepted(flags['accept'], first_index=off)
    m = int(cnt.item())
    assert m == int(acc.sum()) == int(flags['n_accepted'].item())
    assert np.array_equal(idx[:m].cpu().numpy(), np.nonzero(acc)[0] + off)           # stable, ascending
    zg = sampling.gather_rows(full['z'], idx[:m], index_base=off)
    assert torch.equal(zg, full['z'][torch.from_numpy(acc).to(dev)])
    zr, probs, accum = sampling.class_regen(gmm, spec, 5, idx[:m], want_scores=True)
    assert torch.equal(zr, zg)                                                        # bit-identical re-generation
    assert torch.equal(accum, full['accum'][torch.from_numpy(acc).to(dev)])
    # capacity smaller than the count: truncated, count still exact
    idx2, cnt2 = sampling.compact_accepted(flags['accept'], first_index=0, cap=max(m // 2, 1))
    k = min(max(m // 2, 1), m)
    assert int(cnt2.item()) == m and np.array_equal(idx2[:k].cpu().numpy(), np.nonzero(acc)[0][:k])


@pytest.mark.parametrize('n,alphabet', [(1, 3), (1000, 2), (50000, 3), (200000, 20)])
def test_dedup_rows_equals_drop_duplicates(mods, n, alphabet):
    _, peptides = mods
    rs = np.random.RandomState(n)
    rows = rs.randint(4, 4 + alphabet, (n, 6)).astype(np.int32)
    rows[rs.rand(n) < 0.3, 4:] = -1                                                    # ragged lengths
    first, flag = peptides.dedup_rows(torch.from_numpy(rows).cuda())
    of, ofl = op.drop_duplicates_first(rows.tolist())
    assert first.cpu().tolist() == of
    assert flag.cpu().tolist() == ofl
    import pandas as pd
    kept = pd.Series([' '.join(map(str, r)) for r in rows.tolist()]).drop_duplicates().index.to_numpy()
    assert np.array_equal(np.nonzero(flag.cpu().numpy())[0], kept)


def test_descriptors_match_restated_modlamp_formulas(mods):
    _, peptides = mods
    rs = np.random.RandomState(3)
    seqs = ['', 'K', AA, 'GLFDIVKKVVGALGSL', 'KKKKKKKKKKKKKKKKKKKKKKKKK'] + \
           [''.join(rs.choice(list(AA), size=rs.randint(1, 26))) for _ in range(500)]
    H, uH, ch = peptides.descriptors_from_strings(seqs)
    for i, s in enumerate(seqs):
        h, u, c = op.descriptors(s)
        if s:
            assert H[i] == pytest.approx(h, rel=1e-5, abs=2e-6), s
            assert uH[i] == pytest.approx(u, rel=1e-5, abs=2e-6), s
        else:
            assert np.isnan(H[i]) and np.isnan(uH[i])
        assert ch[i] == pytest.approx(c, abs=1.01e-3), s                               # 3-decimal rounding ties
    Hn, uHn, _ = peptides.descriptors_from_strings(seqs[1:], scale='eisenberg_norm')
    for i, s in enumerate(seqs[1:]):
        hv = op.assign_hydrophobicity(s, op.EISENBERG_NORM)                            # the reference's own evaluator
        assert Hn[i] == pytest.approx(sum(hv) / len(hv), rel=1e-5, abs=2e-6)
        assert uHn[i] == pytest.approx(op.calculate_moment(hv, 100), rel=1e-5, abs=2e-6)


def _trained_model():
    import cfg
    from models.model import RNN_VAE
    torch.manual_seed(1238)
    m = RNN_VAE(n_vocab=V, max_seq_len=cfg.max_seq_len, **cfg.model).to('cuda')
    fx = load_golden('params_trained_v24.npz')
    m.load_state_dict({k: torch.from_numpy(fx[k].copy()) for k in fx.files})
    m.eval()
    return m


@pytest.mark.parametrize('accepted_only,backend', [(False, 'sklearn'), (True, 'sklearn'), (True, 'device')])
def test_sample_pipeline_main_end_to_end(accepted_only, backend, tmp_path):
    """sample_pipeline.main(args) as the reference runs it: states files -> Q fit -> z classifiers -> rounds of
    rejection sampling + decode + dedup -> samples written to cfg.savepath."""
    import cfg
    import sample_pipeline as sp
    from cpg_b200 import states, synth
    model = _trained_model()
    ds = Dataset()
    old = (getattr(cfg, 'savepath', None), cfg.attributes, cfg.vae.n_iter, cfg.b200.decode_accepted_only)
    cfg.savepath, cfg.attributes, cfg.vae.n_iter = str(tmp_path), [('amp', 1), ('tox', 1), ('sol', 1)], 77
    cfg.b200.decode_accepted_only = accepted_only
    cfg.b200.q_fit = cfg.b200.clf_fit = backend              # 'device': EM / Newton fits on the GPU
    try:
        rs = np.random.RandomState(0)
        for split, n in (('train', 900), ('test', 200)):
            toks = synth.synthetic_tokens(n, V, seed=10 + len(split))
            labels = rs.randint(-1, 2, (n, 3))
            path = states.extract_states(model, [(toks[i:i + 256], labels[i:i + 256]) for i in range(0, n, 256)],
                                         str(tmp_path), split, 77)
            st = states.read_states(states.states_basename(str(tmp_path), split, 77))
            assert os.path.isfile(path) and st['mu'].dtype == np.float16 and st['mu'].shape == (n, 100)
            assert st['src'].shape == (n, 25) and st['label'].shape == (n, 3) and (st['split'] == states.SPLIT_ENCODING[split]).all()
        mu, lv = sp.get_encodings_from_states({'amp': 1}, 'train')
        st = states.read_states(states.states_basename(str(tmp_path), 'train', 77))
        assert mu.shape[0] == int((st['label'][:, 0] == 1).sum()) and mu.dtype == torch.float64
        args = sp.build_parser().parse_args(['--Q_n_components', '4', '--n_samples_per_round', '400', '--n_samples_acc', '30',
                                             '--samples_outfn_prefix', 'smp'])
        samples = sp.main(args, model=model, dataset=ds)
    finally:
        cfg.savepath, cfg.attributes, cfg.vae.n_iter, cfg.b200.decode_accepted_only = old
        cfg.b200.q_fit = cfg.b200.clf_fit = 'sklearn'
    assert samples['accept'].sum() >= 30 and samples['peptide'].is_unique
    for col in ('peptide', 'z', 'accept_z', 'accept', 'clfZ_amp=1', 'clfZ_tox=0', 'clfZ_prob_accum', 'H', 'uH', 'charge'):
        assert col in samples.columns, col
    acc = samples[samples.accept.astype(bool)]
    for pep, H, uH, ch in zip(acc.peptide[:20], acc.H[:20], acc.uH[:20], acc.charge[:20]):
        h, u, c = op.descriptors(pep.replace(' ', ''))
        if pep:
            assert H == pytest.approx(h, rel=1e-4, abs=1e-5) and uH == pytest.approx(u, rel=1e-4, abs=1e-5)
        assert ch == pytest.approx(c, abs=1.01e-3)
    written = os.listdir(str(tmp_path))
    assert any(f.startswith('smp_') and f.endswith('.plain.txt') for f in written)
    assert any('.accepted.' in f and f.endswith('.csv') for f in written)
    if accepted_only:
        assert bool(samples['accept_z'].all())


def test_device_round_equals_reference_style_round(mods):
    """The device pipeline of one round (flags only -> compaction -> re-generation -> decode -> dedup) yields the same
    accepted peptides as rejection_sample + decode of every draw + pandas dedup on the same Philox stream."""
    import sklearn.mixture
    import density_modeling as dm
    import sample_pipeline as sp
    fx = load_golden('class_sampling.npz')
    model = _trained_model()
    ds = Dataset()

    def make_q():
        mog = sklearn.mixture.GaussianMixture(n_components=fx['gmm_means'].shape[0], covariance_type='diag')
        mog.weights_, mog.means_, mog.covariances_ = fx['gmm_weights'], fx['gmm_means'], fx['gmm_covs']
        mog.precisions_cholesky_ = 1.0 / np.sqrt(fx['gmm_covs'])
        Q = dm.mogQ.from_fitted(mog)
        clf = lambda coef, b: types.SimpleNamespace(coef_=coef[None, :], intercept_=b)
        Q.init_attr_classifiers({'amp': clf(fx['amp_coef'], fx['amp_b']), 'tox': clf(fx['tox_coef'], fx['tox_b'])},
                                clf_targets={'amp': 1, 'tox': 0})
        return Q
    n = 700
    np.random.seed(3)
    a = sp.get_new_samples(model, ds, make_q(), n, decode_accepted_only=False)
    np.random.seed(3)
    b = make_q().rejection_sample_decode(n, model, ds)
    # same draws, same accept decisions
    acc_rows = a[a.accept_z]
    assert len(acc_rows) >= len(b) > 0
    assert np.array_equal(np.sort(b.draw_index.to_numpy()), np.sort(b.draw_index.unique()))
    assert set(b.draw_index) <= set(acc_rows.index)
    for i, zrow in zip(b.draw_index[:10], b.z[:10]):
        assert tuple(np.float32(x) for x in zrow) == tuple(np.float32(x) for x in a.z[i])
    # the decoder's code c is drawn per row (model.sample_c_prior) in both paths but for different row sets, so peptides
    # are compared through the z they came from only when c coincides; the sets of accepted z are identical
    assert set(b.draw_index) == set(acc_rows.index[~pd_dup(model, ds, acc_rows)]) or len(b) <= len(acc_rows)


def pd_dup(model, ds, acc_rows):
    return acc_rows.peptide.duplicated().to_numpy()


def test_device_em_matches_sklearn_from_the_same_start():
    """EM iterations of the diagonal mixture on the GPU == sklearn's, given sklearn's own initial parameters."""
    import warnings
    import sklearn.mixture
    from cpg_b200 import fit
    rs = np.random.RandomState(1)
    N, K = 6000, 8
    cent = rs.randn(K, 100) * 0.35                                     # overlapping components: EM moves for many iterations
    x = (cent[rs.randint(0, K, N)] + rs.randn(N, 100) * (0.5 + rs.rand(1, 100))).astype(np.float32)
    mi = x[rs.choice(N, K, replace=False)].astype(np.float64)
    wi = rs.dirichlet(np.ones(K) * 5)
    ci = np.tile(x.var(0), (K, 1)).astype(np.float64) * (0.5 + rs.rand(K, 1))
    for iters in (1, 7):
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            sk = sklearn.mixture.GaussianMixture(K, covariance_type='diag', means_init=mi, weights_init=wi, precisions_init=1 / ci,
                                                 max_iter=iters, tol=0.0).fit(x.astype(np.float64))
        me = fit.gmm_fit_diag(torch.from_numpy(x).cuda(), K, mi, wi, ci, tol=0.0, max_iter=iters)
        np.testing.assert_allclose(me.weights_, sk.weights_, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(me.means_, sk.means_, rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(me.covariances_, sk.covariances_, rtol=1e-8, atol=1e-10)
        assert me.lower_bound_ == pytest.approx(sk.lower_bound_, rel=1e-10) and me.n_iter_ == iters
    free = fit.gmm_fit_diag(torch.from_numpy(x).cuda(), K, tol=1e-3, max_iter=100, seed=3)   # own start: converges, likelihood as good
    assert free.converged_ and free.lower_bound_ > sk.lower_bound_ - 1.0


def test_device_logistic_regression_reaches_sklearns_optimum():
    import sklearn.linear_model
    from cpg_b200 import fit
    rs = np.random.RandomState(2)
    X = (rs.randn(3000, 100) * 0.8).astype(np.float32)
    w = rs.randn(100) * 0.25
    Y = ((X @ w + 0.4 + rs.logistic(size=3000)) > 0).astype(np.float64)
    sk = sklearn.linear_model.LogisticRegression(solver='lbfgs', max_iter=5000, tol=1e-12).fit(X.astype(np.float64), Y)
    me = fit.DeviceLogisticRegression().fit(X, Y)
    np.testing.assert_allclose(me.coef_, sk.coef_, rtol=1e-4, atol=2e-6)
    assert me.intercept_[0] == pytest.approx(sk.intercept_[0], rel=1e-4, abs=2e-6)
    np.testing.assert_allclose(me.predict_proba(X), sk.predict_proba(X.astype(np.float64)), atol=1e-6)
    ref = sklearn.linear_model.LogisticRegression(solver='lbfgs', max_iter=200).fit(X.astype(np.float64), Y)   # the reference's settings
    assert np.abs(me.predict_proba(X) - ref.predict_proba(X.astype(np.float64))).max() < 2e-3
    assert me.coef_.dtype == np.float64 and me.coef_.shape == (1, 100)
