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


def test_bench_work_model_and_traffic_table():
    """bench.py's roofline inputs: every kernel that can dominate the step has an algorithmic-work entry, the byte counts
    follow the stash layout (planes x hidden x rows x 4 B), and the committed ncu traffic table is readable."""
    import bench
    B, L = 4096, 25
    work = bench.kernel_work(B)
    for name in ('k_gru_fwd_enc_tc', 'k_gru_fwd_dec_tc', 'k_gru_bwd_enc_tc', 'k_gru_bwd_dec_tc', 'k_wgrad_tc_enc',
                 'k_wgrad_tc_dec', 'k_dec_out_tc', 'k_sgemm'):
        assert name in work and work[name][1] > 0, name
    assert work['k_gru_fwd_enc_tc'][1] == 2 * B * L * 5 * 80 * 4          # h + 4 gate planes, two directions
    assert work['k_gru_bwd_enc_tc'][1] == 2 * B * L * 9 * 80 * 4          # 5 planes in, 4 dg planes out
    assert work['k_gru_bwd_dec_tc'][1] == B * L * 10 * 104 * 4            # + dh_out
    assert work['k_gru_fwd_enc_tc'] == work['k_gru_fwd_enc']              # SIMT and tcgen05 flavours move the same bytes
    t = bench.load_traffic('k_gru_bwd_enc_tc')
    assert t is None or (0.5 < t / work['k_gru_bwd_enc_tc'][1] < 1.5)     # ncu DRAM bytes ~ algorithmic bytes
    assert bench.load_traffic('no_such_kernel') is None
