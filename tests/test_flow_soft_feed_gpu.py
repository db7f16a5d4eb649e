"""GPU parity of the 'next' rows of SURVEY 8(f): normalising flows and soft sampling against golden vectors of the
LIVE reference (oracle/gen_flow_golden.py), and the on-device data feed against its host restatement."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu
V = 24


@pytest.mark.parametrize('kind,n', [('planar', 3), ('radial', 2), ('alternating', 5)])
def test_flow_forward_matches_reference_golden(kind, n):
    from models import flow
    fx = load_golden('flow.npz')
    m = flow.build_flow(kind, n, 100).to('cuda')
    m.load_state_dict({k[len(kind) + 7:]: torch.from_numpy(fx[k].copy()) for k in fx.files if k.startswith(kind + '/param/')})
    z = torch.from_numpy(fx[kind + '/z']).cuda()
    zt, loss = m(z, train=True)
    np.testing.assert_allclose(zt.cpu().numpy(), fx[kind + '/z_out'], rtol=1e-5, atol=2e-6)
    assert float(loss) == pytest.approx(float(fx[kind + '/loss']), rel=1e-4, abs=1e-6)
    assert getattr(zt, 'flowed', False)
    # "maintain invertibility" moved the same parameters the reference moved
    for k in fx.files:
        if k.startswith(kind + '/param_after/'):
            got = dict(m.state_dict())[k[len(kind) + 13:]].cpu().numpy()
            np.testing.assert_allclose(got, fx[k], rtol=1e-6, atol=1e-7, err_msg=k)
    z2 = m(z, train=False)
    np.testing.assert_allclose(z2.cpu().numpy(), fx[kind + '/z_out'], rtol=1e-5, atol=2e-6)


def _model(flow=0, flow_type=''):
    import cfg
    from models.model import RNN_VAE
    kw = dict(cfg.model)
    kw.update(flow=flow, flow_type=flow_type)
    m = RNN_VAE(n_vocab=V, max_seq_len=cfg.max_seq_len, **kw).to('cuda')
    pf = load_golden('params_trained_v24.npz')
    m.load_state_dict({k: torch.from_numpy(pf[k].copy()) for k in pf.files}, strict=False)
    m.eval()
    return m


@pytest.mark.parametrize('mode,temp', [('greedy_softmax', 1.0), ('greedy_softmax', 0.6), ('none_softmax', 1.0)])
def test_soft_sampling_matches_reference_golden(mode, temp):
    fx = load_golden('flow.npz')
    m = _model()
    z, c = torch.from_numpy(fx['soft/z']), torch.from_numpy(fx['soft/c'])
    ix, soft = m.sample_G(16, z, c, sample_mode=mode, temp=temp)
    tag = 'soft/%s_t%.1f' % (mode, temp)
    assert np.array_equal(ix.cpu().numpy(), fx[tag + '/ix'])                       # token ids bit-exact
    np.testing.assert_allclose(soft.cpu().numpy(), fx[tag + '/soft'], rtol=1e-4, atol=2e-6)


def test_categorical_softmax_and_flow_in_generate_sentences():
    m = _model(flow=4, flow_type='alternating')
    assert m.use_flow and len(list(m.vae_params())) == 20 + 6 * 4                  # flow parameters join the VAE group
    torch.manual_seed(0)
    np.random.seed(0)
    (ix, soft), z, c_ix = m.generate_sentences(9, sample_mode='categorical_softmax', temp=0.8)
    assert ix.shape[0] == 9 and soft.shape[:2] == ix.shape and soft.shape[2] == V and getattr(z, 'flowed', False)
    s = soft.cpu().numpy()
    t = ix.cpu().numpy()
    assert np.allclose(s[:, 0].argmax(1), 2) and np.allclose(s[:, 0].sum(1), 1)    # one-hot <start>
    for r in range(9):
        e = np.where(t[r] == 3)[0]
        stop = e[0] if e.size else t.shape[1]
        assert np.allclose(s[r, 1:stop].sum(1), 1, atol=1e-5)                      # proper softmaxes before <eos>
        assert np.allclose(s[r, stop:], 0)                                         # zeroed from the <eos> step on
    with pytest.raises(ValueError):
        m(torch.full((2, 25), 4, dtype=torch.int64, device='cuda'))                # forward() with flow raises, as the reference
    with pytest.raises(NotImplementedError):
        m.sample_G(2, torch.zeros(2, 100), torch.eye(2), sample_mode='gumbel_soft')


def test_device_data_feed_draws_weighted_batches():
    from cpg_b200 import feed
    rs = np.random.RandomState(0)
    aa = list('ACDEFGHIKLMNPQRSTVWY')
    seqs = [' '.join(rs.choice(aa, size=rs.randint(1, 40))) for _ in range(300)]
    itos = feed.build_vocab(seqs)
    rows = feed.tokenize_and_pad(seqs, itos, 25)
    assert rows.shape == (300, 25) and (rows[:, 0] == 2).all()
    for r, s in zip(rows, seqs):                                                    # Field(fix_length=25): <= 23 residues
        n = min(len(s.split()), 23)
        assert r[n + 1] == 3 and (r[n + 2:] == 1).all() and [itos[i] for i in r[1:n + 1]] == s.split()[:23]
    w = feed.sample_weights(300, [(np.arange(300) < 30, 20), (np.arange(300) < 100, 10)])
    assert w.sum() == pytest.approx(1.0) and w[0] / w[299] == pytest.approx(20) and w[50] / w[299] == pytest.approx(10)
    f = feed.DeviceDataFeed(rows, w, itos, mbsize=200000, seed=3)
    b = f.next_batch('train_vae', want_index=True)
    idx = f.last_index.cpu().numpy()
    assert b.text.dtype == torch.int64 and b.text.is_cuda and tuple(b.text.shape) == (200000, 25)
    assert np.array_equal(b.text.cpu().numpy(), rows[idx].astype(np.int64))
    freq = np.bincount(idx, minlength=300) / idx.size
    assert np.abs(freq - w).max() < 5 * np.sqrt(w.max() / idx.size)
    b2 = f.next_batch('train_vae', want_index=True)
    assert not np.array_equal(idx, f.last_index.cpu().numpy())                      # next step, new draws
    g = feed.DeviceDataFeed(rows, w, itos, mbsize=200000, seed=3)
    assert torch.equal(g.next_batch('train_vae').text, b.text)                      # same seed / step -> same batch
    assert f.idx2sentence(b.text[0], print_special_tokens=False) == ' '.join(seqs[idx[0]].split()[:23])


def test_train_vae_runs_on_the_device_feed():
    import cfg
    import tb_json_logger
    import train_vae
    from cpg_b200 import feed, synth
    toks = synth.synthetic_tokens(500, V, seed=1).numpy().astype(np.uint8)
    itos = list(feed.SPECIALS) + list('ACDEFGHIKLMNPQRSTVWY')
    ds = feed.DeviceDataFeed(toks, None, itos, mbsize=64, seed=1)
    m = _model()
    m.train()
    cfgv = cfg.Bunch(cfg.vae)
    cfgv.update(cfg.shared)
    cfgv.s_iter, cfgv.n_iter, cfgv.cheaplog_every, cfgv.expsvlog_every = 0, 6, 3, 10 ** 9
    tb_json_logger.configure()
    train_vae.train_vae(cfgv, m, ds)
    vals = tb_json_logger.get_values()
    assert sorted(vals) == [0, 3, 6] and all(np.isfinite(v) for it in vals for v in vals[it].values())
