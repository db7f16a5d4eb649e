"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol include/cpg_b200.h
declares, the flat parameter layout, the reference-compatible config surface, host helpers, and the
"no CPU fallback" contract."""
import ctypes
import importlib
import json
import os
import re
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, PKG, ROOT

LIB = os.path.join(PKG, 'cpg_b200', 'libcpg_b200.so')
HEADER = os.path.join(ROOT, 'include', 'cpg_b200.h')


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(cpg_[a-z0-9_]+)\s*\(', src)))


@pytest.fixture(scope='module')
def lib():
    if not os.path.isfile(LIB):
        import __graft_entry__
        __graft_entry__.build()
    return ctypes.CDLL(LIB)


def test_library_exports_every_declared_symbol(lib):
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), 'missing export ' + n
    assert lib.cpg_abi_version() == 1


def test_binding_covers_the_header(lib):
    from cpg_b200 import _lib
    L = _lib.lib()
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared_symbols()
    assert L.cpg_abi_version() == 1


def test_param_layout_matches_reference_shapes(lib):
    from cpg_b200 import _lib, engine
    offs, sizes, total = _lib.param_layout(24)
    shapes = engine.param_shapes(24)
    assert [int(np.prod(s)) for s in shapes.values()] == sizes
    assert sum(sizes) == 258568                              # unique VAE parameters of the reference at V=24
    assert all(o % 4 == 0 for o in offs) and total >= sum(sizes)
    for (o, n), o2 in zip(zip(offs, sizes), offs[1:] + [total]):
        assert o + n <= o2
    fx = np.load(os.path.join(GOLDEN, 'params_init_v24.npz'))
    for name, shp in shapes.items():
        assert tuple(fx[name].shape) == tuple(shp), name
    rc = lib.cpg_vae_param_layout(64, (ctypes.c_int64 * 19)(), (ctypes.c_int64 * 19)())
    assert rc == -1                                          # CPG_EINVAL: n_vocab > 32


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry fails loudly instead of falling back."""
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from cpg_b200 import CpgLibraryError, engine
    with pytest.raises(CpgLibraryError):
        engine.FlatState(24)
    with pytest.raises(CpgLibraryError):
        engine.latent_stats(torch.zeros(2, 100), torch.zeros(2, 100))
    import cfg
    from models.model import RNN_VAE
    m = RNN_VAE(n_vocab=24, max_seq_len=25, **cfg.model)
    with pytest.raises(CpgLibraryError):
        m(torch.zeros(2, 25, dtype=torch.int64))
    out = ctypes.c_void_p()
    L = ctypes.CDLL(LIB)
    L.cpg_last_error.restype = ctypes.c_char_p
    assert L.cpg_create(ctypes.byref(out), 0) != 0
    assert b'CUDA' in L.cpg_last_error() or b'cuda' in L.cpg_last_error()


def test_product_package_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in re.sub(r'#.*', '', src).replace('"""', ''), os.path.join(dirpath, f)


def _fresh_cfg():
    for m in ('cfg',):
        sys.modules.pop(m, None)
    return importlib.import_module('cfg')


def test_cfg_matches_reference_surface():
    ref = json.load(open(os.path.join(GOLDEN, 'cfg_reference.json')))
    cfg = _fresh_cfg()
    d = {}
    cfg._cfg_import_export(d, cfg, mode='fill_dict')
    ours = {k: v for k, v in d.items() if not k.startswith('b200.')}
    assert ours == {k: v for k, v in ref['defaults'].items()}
    cfg._update_cfg()
    d = {}
    cfg._cfg_import_export(d, cfg, mode='fill_dict')
    assert {k: v for k, v in d.items() if not k.startswith('b200.')} == ref['updated']
    assert [a[0] for a in cfg.attributes] == ref['attributes']
    cfg = _fresh_cfg()
    cfg.tiny = True
    cfg._update_cfg()
    d = {}
    cfg._cfg_import_export(d, cfg, mode='fill_dict')
    assert {k: v for k, v in d.items() if not k.startswith('b200.')} == ref['tiny']
    assert cfg.vae.batch_size == 5 and cfg.vae.n_iter == 100
    _fresh_cfg()


def test_cfg_argparse_roundtrip(tmp_path):
    import argparse
    cfg = _fresh_cfg()
    parser = argparse.ArgumentParser(argument_default=argparse.SUPPRESS)
    cfg._cfg_import_export(parser, cfg, mode='fill_parser')
    args = parser.parse_args(['--vae.batch_size', '4096', '--model.E_args.h_dim', '80', '--runname', 'x',
                              '--losses.wae_mmd.sigma', '3.5'])
    cfg._override_config(args, cfg)
    cfg.savepath_toplevel = str(tmp_path)
    cfg._update_cfg()
    assert cfg.vae.batch_size == 4096 and cfg.losses.wae_mmd.sigma == 3.5 and cfg.runname == 'x'
    cfg._save_config(args, cfg, cfg.savepath)
    done = json.load(open(os.path.join(cfg.savepath, 'config_complete.json')))
    assert done['vae.batch_size'] == 4096 and done['b200.fused_step'] is True
    assert json.load(open(os.path.join(cfg.savepath, 'config_overrides.json')))['runname'] == 'x'
    _fresh_cfg()


def test_anneal_and_beam_helper():
    import cfg
    import utils
    from oracle import wae as ow
    for it in (0, 1, 17, 39999, 40000, 123456):
        assert utils.anneal(cfg.vae.beta, it) == ow.anneal_beta(it)
    from models.Beam import Beam
    b = Beam(3, pad=1, bos=2, eos=3, n_best=2)
    lp = torch.log_softmax(torch.tensor([[0.1, 0.0, 0.0, 2.0, 1.0]] * 3), 1)
    b.advance(lp.clone())
    assert b.get_current_state().tolist() == [3, 4, 0] and b.eos_top and len(b.finished) == 1
    assert [int(t) for t in b.get_hyp(1, 0)] == [2, 3]


def test_shard_bounds():
    from cpg_b200.parallel import shard_bounds
    for n, w in ((10, 3), (4096, 8), (5, 8), (0, 2)):
        spans = [shard_bounds(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def test_bench_roofline_model_is_the_algorithmic_one():
    """the roofline is taken against ALGORITHMIC bytes (h-stash once each way, 60 KB/seq for the
    step), not the bytes the implementation chooses to move; the dominant kernel is the largest share whatever
    it is; dram traffic and traffic_ratio are reported next to it."""
    import bench
    B = 4096
    work = bench.kernel_work(B)
    assert work['k_gru_bwd_enc'][1] == 2 * B * 25 * 80 * 4            # h stash of both directions, read once
    assert work['k_gru_fwd_dec'][1] == B * 25 * 102 * 4 + B * 25
    sw = bench.step_work(B)
    assert 55e3 * B < sw['alg_bytes'] - 6.2e6 < 65e3 * B               # ~60 KB/seq
    assert bench.work_for('k_gemm_tc[4096x500x100]', work, B)[0] == 2 * 4096 * 500 * 100
    assert bench.work_for('k_gru_bwd_enc_tc', work, B) == work['k_gru_bwd_enc']
    peaks = {'hbm_gbs': 6537.0, 'bf16_tflops_sustained': 1407.2}
    rows = [('k_gru_bwd_enc_tc', 0.120 * 4, 4), ('k_gemm_tc[4096x500x100]', 0.9 * 4, 12), ('k_clip_adam', 0.01 * 4, 4)]
    r = bench.build_roofline(rows, 4, B, 1.0, peaks, 'measured')
    assert r['kernel'] == 'k_gemm_tc[4096x500x100]' and r['bound'] == 'tensor'    # largest share wins, no exclusions
    pk = r['per_kernel']['k_gru_bwd_enc_tc']
    assert pk['frac_hbm_algorithmic'] == pytest.approx(work['k_gru_bwd_enc'][1] / 120e-6 / 6537e9, rel=1e-2)
    assert r['step']['frac_hbm_algorithmic'] == pytest.approx(sw['alg_bytes'] / 1e-3 / 6537e9, rel=1e-2)
    for k in ('frac', 'frac_algorithmic', 'frac_dram', 'traffic', 'traffic_ratio', 'achieved', 'peak', 'unit', 'step'):
        assert k in r


def test_package_synthetic_workload_equals_oracle_generator():
    """bench.py's repo arm builds its inputs with the package (no oracle import on the timed path); the
    CPU-baseline arm uses the oracle's generator: both must yield the same workload."""
    from cpg_b200 import synth
    from oracle import cpu_baseline as cb
    from oracle import wae as ow
    for B, seed in ((7, 3), (512, 1238)):
        assert torch.equal(synth.synthetic_tokens(B, 24, seed), ow.synthetic_tokens(B, 24, seed))
    a, b = synth.synthetic_class_setup(), cb.synthetic_class_setup()
    assert all(np.array_equal(x, y) for x, y in zip(a[:3], b[:3]))
    assert all(np.array_equal(x[1], y[1]) and x[3] == y[3] for x, y in zip(a[3], b[3]))


def test_states_files_round_trip_float16(tmp_path):
    """states_{split}_{n_iter}: same datasets / dtypes as vis/scripts/build_index.py:30-66 (float16 encodings)."""
    from cpg_b200 import states
    rs = np.random.RandomState(0)
    n = 37
    mu, lv = rs.randn(n, 100).astype(np.float32), rs.randn(n, 100).astype(np.float32) - 2
    src = rs.randint(0, 24, (n, 25))
    lab = rs.randint(-1, 2, (n, 3))
    base = states.states_basename(str(tmp_path), 'val', 5)
    path = states.write_states(base, src, mu, mu, lv, lab, np.full((n, 1), 1))
    assert os.path.isfile(path)
    st = states.read_states(base)
    assert set(st) == set(states.KEYS)
    assert st['mu'].dtype == np.float16 and st['logvar'].dtype == np.float16 and st['z'].dtype == np.float16
    assert np.array_equal(st['mu'], mu.astype(np.float16)) and np.array_equal(st['src'], src)
    assert st['label'].dtype.kind == 'i' and st['split'].shape == (n, 1)
    with pytest.raises(FileNotFoundError):
        states.read_states(states.states_basename(str(tmp_path), 'test', 5))


def test_sample_pipeline_surface_matches_the_reference_names():
    """Function names / CLI flags of the reference's sample_pipeline.py (:41-361) and api.py are present."""
    import inspect
    sp = importlib.import_module('sample_pipeline')
    for name, params in (('get_encodings', ['query', 'split', 'model', 'dataloader']),
                         ('get_encodings_from_dataloader', ['query', 'split', 'model', 'dataloader']),
                         ('get_encodings_from_states', ['query', 'split']),
                         ('fitQ_and_test', ['QClass', 'QKwargs', 'Q_select', 'negative_select', 'model', 'dataloader']),
                         ('decode_from_z', ['z', 'model', 'dataset']), ('save_csv_pkl', ['samples', 'fn']),
                         ('save_samples', ['samples', 'basedir', 'fn_prefix']), ('score_clfZ', ['clf', 'z']),
                         ('build_clfZ', ['attr']), ('get_new_samples', ['model', 'dataset', 'Q', 'n_samples']),
                         ('compute_modlamp', ['df']), ('one_sampling_round', ['model', 'dataset', 'Q', 'n_samples_per_round']),
                         ('get_sample_source_str', []), ('main', ['args'])):
        fn = getattr(sp, name)
        got = list(inspect.signature(fn).parameters)
        assert got[:len(params)] == params, (name, got)
    flags = {a.option_strings[0] for a in sp.build_parser()._actions if a.option_strings}
    for f in ('--QClass', '--Q_n_components', '--Q_covariance_type', '--n_samples_per_round', '--n_samples_acc',
              '--samples_outfn_prefix', '--Q_select_amppos', '--Q_from_full_dataloader', '--vae.batch_size', '--seed'):
        assert f in flags, f
    assert sp.Q_CLASS.__name__ == 'mogQ' and set(sp.Q_KWARGS) == {'n_components', 'z_num_samples', 'covariance_type'}
    api = importlib.import_module('api')
    for name in ('Vocab', 'load_trained_model', 'encode_sequence', 'sample_from_model', 'get_model_and_vocab_path',
                 'get_result_for_model', 'main'):
        assert hasattr(api, name), name


def test_vocab_and_peptide_formula_known_answers(tmp_path):
    api = importlib.import_module('api')
    fn = tmp_path / 'vocab.dict'
    words = ['<unk>', '<pad>', '<start>', '<eos>'] + list('ACDEFGHIKLMNPQRSTVWY')
    fn.write_text(''.join('%s %d\n' % (w, i) for i, w in enumerate(words)))
    v = api.Vocab(str(fn))
    assert v.size() == 24 and v.special_tokens_ix == {0, 1, 2, 3}
    ix = v.to_ix('K L A')
    assert ix.shape == (1, 25) and ix[0, :5].tolist() == [2, 12, 13, 4, 3] and int(ix[0, -1]) == 1
    assert v.to_word(ix[0], print_special_tokens=False) == ['K', 'L', 'A']
    from oracle import peptides as op
    # hand-computed: one lysine, amidated C-terminus, pH 7
    q = 10 ** 9.38 / (10 ** 9.38 + 1e7) + 10 ** 10.67 / (10 ** 10.67 + 1e7) - 1e7 / (1e15 + 1e7)
    assert op.charge('K') == round(q, 3) == 1.996
    assert op.descriptors('K')[:2] == (-1.5, 1.5)
    h = [0.62, 1.06]                                             # 'AL': moment of two residues 100 degrees apart
    want = (((h[0] + h[1] * np.cos(np.deg2rad(100))) ** 2 + (h[1] * np.sin(np.deg2rad(100))) ** 2) ** 0.5) / 2
    assert op.descriptors('AL')[1] == pytest.approx(want, rel=1e-12)
    assert op.drop_duplicates_first(['a', 'b', 'a', 'c', 'b']) == ([0, 1, 0, 3, 1], [1, 1, 0, 1, 0])


def test_rejection_sample_host_copy_ring_logic(monkeypatch):
    """density_modeling._to_host_chunked: chunks through a two-slot ring into fresh host tensors, every dtype the
    sampler returns (fp32 z, fp64 / fp32 scores, bool mask), sizes below / across / at multiples of the ring size.
    The CUDA pieces (pinned allocation, stream events) are replaced by host stand-ins: this checks the indexing."""
    sys.path.insert(0, PKG)
    import types
    dm = importlib.import_module('density_modeling')

    class Ev:
        def synchronize(self):
            pass

    class St:
        def record_event(self):
            return Ev()

    monkeypatch.setattr(dm, 'RING_BYTES', 1000)
    monkeypatch.setitem(dm._RING, 0, (torch.empty(2, 1000, dtype=torch.uint8), [None, None]))
    monkeypatch.setattr(torch.cuda, 'current_stream', lambda dev=None: St())
    dev = types.SimpleNamespace(index=0)
    g = torch.Generator().manual_seed(3)
    ts = [torch.randn(37, 100, generator=g), torch.randn(3, 37, generator=g, dtype=torch.float64),
          torch.randn(250, generator=g), torch.rand(37, generator=g) > 0.5, torch.zeros(0, 5), torch.randn(125, 2, generator=g)]
    outs = dm._to_host_chunked(ts, dev)
    for a, b in zip(ts, outs):
        assert a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b)
        assert b.data_ptr() != a.data_ptr() or a.numel() == 0


def test_rows_to_sentences_equals_the_dataset_loop():
    """peptides.rows_to_sentences (the round table's peptide strings without a Python loop over tokens) against the
    reference's idx2sentences semantics (data_processing/dataset.py:288-300: special ids dropped, words joined by ' ')."""
    sys.path.insert(0, PKG)
    from cpg_b200 import peptides
    words = ['<unk>', '<pad>', '<start>', '<eos>'] + list('ACDEFGHIKLMNPQRSTVWY')
    ds = type('D', (), {'idx2sentences': staticmethod(
        lambda seqs, print_special_tokens=True: [' '.join(words[int(i)] for i in s if print_special_tokens or int(i) > 3) for s in seqs])})
    assert peptides.vocabulary_words(ds, 24) == words
    rng = np.random.default_rng(5)
    for lo in (0, 4):                                   # with / without special ids inside the kept part of a row
        tok = rng.integers(lo, 24, size=(5000, 25)).astype(np.int32)
        ln = rng.integers(0, 26, size=5000)
        tok[3] = 1
        ln[4] = 0
        want = ds.idx2sentences([r[:k] for r, k in zip(tok.tolist(), ln.tolist())], print_special_tokens=False)
        assert peptides.rows_to_sentences(tok, ln, words) == want
    # vocabularies with longer words take the caller's fallback (or the plain loop)
    w2 = words[:4] + ['Ala', 'Gly']
    assert peptides.rows_to_sentences(np.array([[4, 5, 3, 5]]), [4], w2) == ['Ala Gly Gly']
    assert peptides.rows_to_sentences(np.array([[4, 5, 3, 5]]), [3], w2, fallback=lambda rows: ['x%d' % len(r) for r in rows]) == ['x3']
    assert peptides.rows_to_sentences(np.zeros((0, 25), np.int32), [], words) == []
