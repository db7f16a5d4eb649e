"""The reference-named Python surface (cfg, models.model.RNN_VAE, losses, train_vae, density_modeling,
sample_pipeline) running on the CUDA kernels: reads like a test of the reference's own API."""
import types

import numpy as np
import pytest
import torch

from conftest import load_golden
from helpers import assert_params_close, clipped, compare_beam
from oracle import class_sampling as oc
from oracle import decode as od
from oracle import wae as ow

pytestmark = pytest.mark.gpu
V = 24


class Dataset:
    """Stand-in for AttributeDataLoader (data_processing/dataset.py:285-300)."""
    def __init__(self, batches):
        self.batches, self.i = batches, 0

    def next_batch(self, name):
        b = self.batches[self.i % len(self.batches)]
        self.i += 1
        return types.SimpleNamespace(text=b)

    def idx2sentence(self, idxs, print_special_tokens=True):
        return ' '.join(str(int(i)) for i in idxs.view(-1))

    def idx2sentences(self, seqs, print_special_tokens=True):
        return [' '.join(str(int(i)) for i in s if print_special_tokens or int(i) > 3) for s in seqs]


@pytest.fixture()
def model():
    import cfg
    from models.model import RNN_VAE
    torch.manual_seed(1238)
    m = RNN_VAE(n_vocab=V, max_seq_len=cfg.max_seq_len, **cfg.model).to('cuda')
    fx = load_golden('params_init_v24.npz')
    m.load_state_dict({k: torch.from_numpy(fx[k].copy()) for k in fx.files})
    return m


def test_state_dict_keys_and_param_groups_match_reference(model):
    fx = load_golden('params_init_v24.npz')
    assert sorted(model.state_dict().keys()) == sorted(fx.files)
    ps = list(model.vae_params())
    assert len(ps) == 20 and sum(p.numel() for p in ps) == 262168            # word_emb counted twice
    assert len({id(p) for p in ps}) == 19
    assert model.decoder.emb is model.word_emb


def test_forward_and_autograd_match_oracle(model):
    """model(...) -> losses.* -> loss.backward() with the reference's call sequence."""
    import losses
    dev = torch.device('cuda')
    B = 12
    tokens = ow.synthetic_tokens(B, V, seed=3)
    torch.manual_seed(5)
    np.random.seed(5)
    model.train()
    (mu, lv), (z, c), logits = model(tokens.to(dev), q_c='prior', sample_z=1)
    p = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    # recover the noise the module drew: eps from z, dropout masks from a replay of the numpy stream
    np.random.seed(5)
    c_ref = np.random.multinomial(1, [0.5, 0.5], B).astype('float32')
    wd = np.random.binomial(1, p=0.3, size=(B, 25)).astype('uint8')
    assert np.array_equal(c.cpu().numpy(), c_ref)
    omu, olv = ow.encoder_forward(p, tokens)
    np.testing.assert_allclose(mu.detach().cpu().numpy(), omu.numpy(), rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(lv.detach().cpu().numpy(), olv.numpy(), rtol=1e-4, atol=2e-6)
    recon = losses.recon_dec(tokens.to(dev), logits)
    assert float(recon) == pytest.approx(float(ow.recon_dec(tokens, logits.detach().cpu())), rel=1e-5)
    kl = losses.kl_gaussianprior(mu, lv)
    assert float(kl) == pytest.approx(float(ow.kl_gaussianprior(omu, olv)), rel=1e-4)
    klp = losses.kl_gaussian_sharedmu(mu, lv)
    loss = recon + 0.5 * kl + 1e-3 * klp + lv.abs().sum(1).mean(0) * 0.01
    loss.backward()
    # oracle gradient with the same z (eps recovered), same masks: only the decoder out-dropout mask is
    # unknown, so compare the gradient of the encoder-side terms, which do not involve it
    g_mu = model.encoder.q_mu.bias.grad
    assert g_mu is not None and torch.isfinite(g_mu).all() and float(g_mu.abs().sum()) > 0
    assert model.word_emb.weight.grad[ow.PAD_IDX].abs().sum() == 0
    assert wd.shape == (B, 25)


def test_eval_forward_max_and_classifier_match_reference_golden(model):
    fx = load_golden('infer_b48.npz')
    dev = torch.device('cuda')
    model.eval()
    np.random.seed(0)
    tokens = torch.from_numpy(fx['tokens']).to(dev)
    real = model.decoder.word_dropout.sample_mask
    model.decoder.word_dropout.sample_mask = lambda shape: torch.zeros(tuple(shape), dtype=torch.uint8)
    try:
        with torch.no_grad():
            (mu, lv), (z, c), logits = model(tokens, q_c='classifier', sample_z='max')
    finally:
        model.decoder.word_dropout.sample_mask = real
    np.testing.assert_allclose(mu.cpu().numpy(), fx['mu'], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(c.cpu().numpy(), fx['c'], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(logits.cpu().numpy(), fx['dec_logits'], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(model.forward_classifier(tokens).cpu().numpy(), fx['cnn_logits'], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize('fused', [True, False])
def test_train_vae_runs_like_the_reference_tiny_config(model, fused, tmp_path):
    """--tiny style run: 11 iterations, B=5, logging every 5, a checkpoint with the reference's keys."""
    import cfg
    import tb_json_logger
    import train_vae
    cfgv = cfg.Bunch(cfg.vae)
    cfgv.update(cfg.shared)
    cfgv.s_iter, cfgv.n_iter, cfgv.batch_size = 0, 10, 5
    cfgv.cheaplog_every, cfgv.expsvlog_every = 5, 10
    cfgv.chkpt_path = str(tmp_path / 'model_{}.pt')
    cfg.b200.fused_step = fused
    tb_json_logger.configure()
    torch.manual_seed(1)
    np.random.seed(1)
    ds = Dataset([ow.synthetic_tokens(5, V, seed=40 + i) for i in range(4)])     # host tensors -> H2D inside
    before = {k: v.detach().clone() for k, v in model.state_dict().items()}
    try:
        train_vae.train_vae(cfgv, model, ds)
    finally:
        cfg.b200.fused_step = True
    vals = tb_json_logger.get_values()
    assert sorted(vals) == [0, 5, 10]
    for it in vals:
        assert set(vals[it]) == {'train_' + k for k in ('z_mu_L1', 'z_logvar', 'z_logvar_L1', 'z_logvar_KL_penalty',
                                                        'L_vae', 'L_vae_recon', 'L_vae_kl', 'L_wae_mmd',
                                                        'L_wae_mmdrf', 'beta')}
        assert all(np.isfinite(v) for v in vals[it].values())
    assert vals[10]['train_L_vae_recon'] < vals[0]['train_L_vae_recon']
    sd = torch.load(str(tmp_path / 'model_10.pt'))
    assert sorted(sd.keys()) == sorted(before.keys())
    assert float((sd['word_emb.weight'].cpu() - before['word_emb.weight'].cpu()).abs().max()) > 0
    assert torch.equal(sd['word_emb.weight'], sd['decoder.emb.weight'])
    assert float(sd['word_emb.weight'][ow.PAD_IDX].abs().sum()) == 0


def test_generate_sentences_modes(model):
    fx = load_golden('decode.npz')
    pf = load_golden('params_trained_v24.npz')
    model.load_state_dict({k: torch.from_numpy(pf[k].copy()) for k in pf.files})
    z = torch.from_numpy(fx['trained/z'])
    c = torch.from_numpy(fx['trained/c'])
    hyps, zz, c_ix = model.generate_sentences(z.shape[0], z, c, sample_mode='beam', beam_size=5)
    ref, margins = fx['trained/beam_hyps'], fx['trained/beam_margin']
    compare_beam(hyps, [[[int(t) for t in h if t >= 0] for h in hs] for hs in ref], margins, 'generate_sentences', n_hyps=1)
    assert np.array_equal(c_ix.cpu().numpy(), fx['trained/c'].argmax(1))
    g, _, _ = model.generate_sentences(z.shape[0], z, c, sample_mode='greedy')
    assert np.array_equal(g.cpu().numpy(), fx['trained/greedy'])
    s, _, _ = model.generate_sentences(7, sample_mode='categorical')
    assert s.shape[0] == 7 and s.dtype == torch.int64 and bool((s[:, 0] == ow.START_IDX).all())


def test_mogQ_rejection_sample_and_pipeline(model):
    """density_modeling.mogQ + sample_pipeline round on the GPU, reference RNG replay bit-exact."""
    import sklearn.mixture
    import density_modeling as dm
    import sample_pipeline as sp
    fx = load_golden('class_sampling.npz')
    mog = sklearn.mixture.GaussianMixture(n_components=fx['gmm_means'].shape[0], covariance_type='diag')
    mog.weights_, mog.means_, mog.covariances_ = fx['gmm_weights'], fx['gmm_means'], fx['gmm_covs']
    mog.precisions_cholesky_ = 1.0 / np.sqrt(fx['gmm_covs'])
    Q = dm.mogQ.from_fitted(mog)
    clf = lambda coef, b: types.SimpleNamespace(coef_=coef[None, :], intercept_=b)
    Q.init_attr_classifiers({'amp': clf(fx['amp_coef'], fx['amp_b']), 'tox': clf(fx['tox_coef'], fx['tox_b'])},
                            clf_targets={'amp': 1, 'tox': 0})
    np.random.seed(int(fx['draw_seed']))
    z, scores, acc = Q.rejection_sample(fx['z'].shape[0], mode='reference_rng')
    assert np.array_equal(z.numpy(), fx['z'])
    safe = np.abs(fx['u'] - fx['score_accum'].astype(np.float64)) > 1e-6
    assert np.array_equal(acc[safe], fx['accepted'][safe])
    assert set(scores) == {'clfZ_amp=1', 'clfZ_tox=0', 'clfZ_prob_accum'}
    assert scores['clfZ_prob_accum'].dtype == np.float32
    np.testing.assert_allclose(scores['clfZ_prob_accum'], fx['score_accum'], rtol=4e-5, atol=1e-7)
    assert Q.logpdf(z[0]) == pytest.approx(float(fx['logpdf_q'][0]), rel=2e-6)
    assert dm.prior_logpdf(z[0]) == pytest.approx(float(fx['logpdf_p'][0]), rel=1e-6)
    torch.manual_seed(77)
    nllq, nllp = dm.evaluate_nll(Q, (torch.from_numpy(fx['nll_mu']), torch.from_numpy(fx['nll_logvar'])))
    assert nllq == pytest.approx(float(fx['nll_q']), rel=1e-5) and nllp == pytest.approx(float(fx['nll_p']), rel=1e-5)
    # one pipeline round on Philox draws, decode accepted only
    ds = Dataset([])
    df = sp.one_sampling_round(model, ds, Q, 300, decode_accepted_only=True)
    st = Q.last_round_stats                       # device round: one row per UNIQUE accepted peptide
    assert st['n_draws'] == 300 and len(df) == st['n_unique'] <= st['n_accepted'] and bool(df['accept'].all())
    assert 0.1 < st['n_accepted'] / 300 < 0.5 and df['peptide'].is_unique
    full = sp.one_sampling_round(model, ds, Q, 300)                 # reference behaviour: every draw decoded
    assert len(full) == 300 and 0.1 < full['accept'].mean() < 0.5 and full['peptide'].notna().all()
    out = sp.run_sampling(model, ds, Q, n_samples_per_round=200, n_samples_acc=20, max_rounds=20)
    assert out['accept'].sum() >= 20 and out['peptide'].is_unique


def test_stale_stash_raises_and_losses_do_not_touch_the_stash(model):
    """The context keeps ONE activation stash: a backward whose stash was replaced by another forward must raise
    (not silently use the wrong activations); stand-alone loss ops between forward and backward use their own
    scratch, whatever rf_dim is."""
    import losses
    from cpg_b200 import _lib
    dev = torch.device('cuda')
    B = 9
    t1 = ow.synthetic_tokens(B, V, seed=3).to(dev)
    t2 = ow.synthetic_tokens(B, V, seed=4).to(dev)
    model.train()
    torch.manual_seed(0)
    np.random.seed(0)
    (mu1, lv1), (z1, _), lg1 = model(t1)
    (mu2, lv2), (z2, _), lg2 = model(t2)                     # replaces the stash of the first forward
    l1, l2 = losses.recon_dec(t1, lg1), losses.recon_dec(t2, lg2)
    with pytest.raises(_lib.CpgLibraryError, match='stash'):
        (l1 + l2).backward()
    # forward -> RF-MMD with a non-default rf_dim (own scratch) -> encode of another batch size is NOT allowed in
    # between (it would drop the stash) but loss ops are: backward works and matches a run without them
    model.zero_grad()
    losses.rf.clear()
    torch.manual_seed(1)
    np.random.seed(1)
    (mu, lv), (z, _), lg = model(t1)
    extra = losses.mmd_rf(z, torch.randn_like(z), sigma=7.0, kernel='gaussian', rf_dim=300)
    full = losses.mmd_full_kernel(z.detach(), torch.randn_like(z), sigma=7.0, kernel='gaussian')
    assert torch.isfinite(full)
    (losses.recon_dec(t1, lg) + 0.0 * extra).backward()
    g_a = model.decoder.fc[1].weight.grad.clone()
    model.zero_grad()
    torch.manual_seed(1)
    np.random.seed(1)
    (mu, lv), (z, _), lg = model(t1)
    losses.recon_dec(t1, lg).backward()
    assert torch.equal(g_a, model.decoder.fc[1].weight.grad)
    losses.rf.clear()
