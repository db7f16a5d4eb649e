"""Module-as-config with the reference's attribute names (cfg.py of
IBM/controlled-peptide-generation): every public module attribute of type float/str/int/bool,
recursively through `Bunch`es, is a `--a.b.c` command-line flag and a key of config_complete.json
(reference cfg.py:56-72); `_update_cfg()` post-processes paths, --tiny, part/partN job splitting and
seeds (reference cfg.py:75-136).  New keys for the B200 engine live under `b200`.

The values are the reference defaults; the file is written around a declarative table instead of
assignment statements, but `cfg.vae.batch_size`, `cfg.model.E_args.h_dim`, ... resolve identically.
"""
import json
import os

from utils import check_dir_exists


class Bunch(dict):
    """dict whose items are also attributes (reference cfg.py:8-11)."""

    def __init__(self, *args, **kwds):
        super().__init__(*args, **kwds)
        self.__dict__ = self


def _bunchify(obj):
    if isinstance(obj, dict) and not isinstance(obj, Bunch):
        return Bunch((k, _bunchify(v)) for k, v in obj.items())
    return obj


_SCALARS = (float, str, int, bool)


def _walk(node):
    """Yield (name, value) for the public entries of a module or Bunch, sorted like dir()."""
    names = sorted(node.keys()) if isinstance(node, dict) else dir(node)
    for name in names:
        if not name.startswith('_'):
            yield name, (node[name] if isinstance(node, dict) else getattr(node, name))


def _cfg_import_export(cfg_interactor, cfg_, prefix='', mode='fill_parser'):
    """fill_parser: add --prefix.key flags; fill_dict: export; override: import present keys."""
    for key, val in _walk(cfg_):
        flat = prefix + key
        if type(val) in _SCALARS:
            if mode == 'fill_parser':
                cfg_interactor.add_argument('--' + flat, type=type(val), help='default: {}'.format(val))
            elif mode == 'fill_dict':
                cfg_interactor[flat] = val
            elif mode == 'override':
                if flat in cfg_interactor:
                    new = getattr(cfg_interactor, flat) if not isinstance(cfg_interactor, dict) else cfg_interactor[flat]
                    if isinstance(cfg_, dict):
                        cfg_[key] = new
                    else:
                        setattr(cfg_, key, new)
            else:
                raise ValueError('unknown mode ' + mode)
        elif type(val) is Bunch:
            _cfg_import_export(cfg_interactor, val, prefix=flat + '.', mode=mode)


def _override_config(args, cfg):
    """Import overrides from an argparse namespace (only the flags that were given)."""
    _cfg_import_export(args, cfg, mode='override')


def _override_config_from_json(cfg, config_json):
    if config_json:
        with open(config_json) as fh:
            _cfg_import_export(Bunch(json.load(fh)), cfg, mode='override')


def _copy_to_nested_dict(cfg_):
    out = {}
    for key, val in _walk(cfg_):
        if type(val) in _SCALARS:
            out[key] = val
        elif type(val) is Bunch:
            out[key] = _copy_to_nested_dict(val)
    return out


def _save_config(cfg_overrides, cfg_complete, savepath):
    fn = os.path.join(savepath, 'config_overrides.json')
    check_dir_exists(fn)
    with open(fn, 'w') as fh:
        json.dump(vars(cfg_overrides), fh, indent=2, sort_keys=True)
    flat = {}
    _cfg_import_export(flat, cfg_complete, mode='fill_dict')
    with open(os.path.join(savepath, 'config_complete.json'), 'w') as fh:
        json.dump(flat, fh, indent=2, sort_keys=True)


def _print(cfg_, prefix=''):
    for key, val in _walk(cfg_):
        if type(val) in _SCALARS:
            print('{}{}\t{}'.format(prefix, key, val))
        elif type(val) is Bunch:
            print('{}{}:'.format(prefix, key))
            _print(val, prefix + '  |- ')


# ------------------------------------------------------------------------------ default values
_VAE_ITERS = 200000
_DEFAULTS = {
    # general
    'config_json': '', 'ignore_gpu': False, 'seed': 1238, 'tiny': False,
    # paths
    'tb_toplevel': 'tb', 'savepath_toplevel': 'output', 'runname': 'default', 'datapath': 'data',
    'loadpath': 'auto', 'vocab_path': 'auto',
    'phase': -1, 'part': 0, 'partN': 1, 'resume_result_json': True,
    # phase 1: VAE / WAE pre-training
    'vae': {
        'batch_size': 32, 'lr': 1e-3, 's_iter': 0, 'n_iter': _VAE_ITERS,
        'beta': {'start': {'val': 1.0, 'iter': 0}, 'end': {'val': 2.0, 'iter': _VAE_ITERS // 5}},
        'lambda_logvar_L1': 0.0, 'lambda_logvar_KL': 1e-3,
        'z_regu_loss': 'mmdrf',          # kl | mmd | mmdrf
        'cheaplog_every': 500, 'expsvlog_every': 20000,
    },
    # phase 2: full training (controlled generation)
    'full': {
        'batch_size': 32, 'lrE': 3e-4, 'lrG': 3e-4, 'lrC': 3e-4,
        'n_iter': 50000, 's_iter': _VAE_ITERS, 'classifier_min_length': 5,
        'beta': {'start': {'val': 2.0, 'iter': _VAE_ITERS}, 'end': {'val': 2.0, 'iter': _VAE_ITERS + 50000}},
        'z_regu_loss': 'mmdrf',
        'C_hard_sample_kwargs': {'sample_mode': 'categorical'},
        'G_soft_sample_kwargs': {'sample_mode': 'none_softmax'},
        'softmax_temp': {'start': {'iter': _VAE_ITERS, 'val': 1.0}, 'end': {'iter': _VAE_ITERS + 50000, 'val': 1.0}},
        'lambda_e': 0.1, 'lambda_c': 1.0, 'lambda_z': 0.1, 'lambda_u': 0.1,
        'lambda_logvar_L1': 0.0, 'lambda_logvar_KL': 1e-3,
        'cheaplog_every': 50, 'expsvlog_every': 2000,
    },
    'shared': {'clip_grad': 5.0},
    'evals': {'sample_size': 2000, 'sample_modes': {'beam': {'sample_mode': 'beam', 'beam_size': 5, 'n_best': 3}}},
    'losses': {'wae_mmd': {'sigma': 7.0, 'kernel': 'gaussian', 'rf_dim': 500, 'rf_resample': False}},
    'max_seq_len': 25,
    'model': {
        'z_dim': 100, 'c_dim': 2, 'emb_dim': 150, 'pretrained_emb': None, 'freeze_embeddings': False,
        'flow': 0, 'flow_type': '',
        'E_args': {'h_dim': 80, 'biGRU': True, 'layers': 1, 'p_dropout': 0.0},
        'G_args': {
            'G_class': 'gru',
            'GRU_args': {'p_word_dropout': 0.3, 'p_out_dropout': 0.3, 'skip_connetions': False},
            'deconv_args': {'max_seq_len': 25, 'num_filters': 100, 'kernel_size': 4, 'num_deconv_layers': 3,
                            'useRNN': False, 'temperature': 1.0, 'use_batch_norm': True, 'num_conv_layers': 2,
                            'add_final_conv_layer': True},
        },
        'C_args': {'min_filter_width': 3, 'max_filter_width': 5, 'num_filters': 100, 'dropout': 0.5},
    },
    # B200 engine switches (new; exported like every other key)
    'b200': {
        'fused_step': True,        # train_vae() runs the fused C-ABI iteration instead of autograd + torch.optim
        'noise_seed': 1238,        # Philox key of the perf-mode noise
        'full_mmd_every': 1,       # the reference evaluates the (log-only) full-kernel MMD every iteration
        'sync_scalars_every': 0,   # > 0: also read the 16-float scalar block back every n-th iteration
                                   # (train_vae.last_scalars; e.g. for a NaN watchdog); 0 = log iterations only
        'decode_accepted_only': False,   # sample_pipeline: True = compact the accepted z on the device and decode only those
                                         # (rows of the round table = unique accepted peptides); False = decode every draw
        'q_fit': 'sklearn',        # mogQ fit: 'sklearn' (reference: host GaussianMixture.fit) | 'device' (EM kernels, cpg_b200.fit)
        'clf_fit': 'sklearn',      # z-space classifiers: 'sklearn' (LogisticRegression lbfgs) | 'device' (Newton, GPU statistics)
        'dp_graph': True,          # under torch.distributed: replay the iteration (collectives included) from one captured CUDA graph per rank
        'dp_full_mmd': 'local',    # under torch.distributed: 'local' = the log-only full-kernel MMD of this rank's shard,
                                   # 'global' = all-gather z / z_prior and evaluate the global-batch value
    },
    'dataset': 'amp',
}
for _k, _v in _DEFAULTS.items():
    globals()[_k] = _bunchify(_v)
del _k, _v
# the full phase continues where the VAE phase stops (reference cfg.py:189-232)
full.beta.start.iter = full.s_iter
full.beta.end.iter = full.s_iter + full.n_iter
full.softmax_temp.start.iter = full.s_iter
full.softmax_temp.end.iter = full.s_iter + full.n_iter

data_kwargs, data_prefixes, attributes = None, None, None      # filled by _set_dataset()

DATA_ROOT = './PATH_TO_DATA/'
amp_sample_prob_factors = {
    'amp=amp_posc': 20, 'amp=amp_posnc': 10, 'amp=amp_negc': 20, 'amp=amp_negnc': 10,
    'tox=tox_posc': 20, 'tox=tox_posnc': 10, 'tox=tox_negc': 20, 'tox=tox_negnc': 10,
    'sol': 20, 'anticancer': 20, 'antihyper': 20, 'hormone': 20,
}


def _weighted(subset):
    return Bunch(subset=subset, weighted_random_sample=True, sample_prob_factors=amp_sample_prob_factors)


amp = Bunch(
    data_kwargs=Bunch(
        lower=False,
        data_path=os.environ.get('DATA_PATH_AMP', DATA_ROOT + 'amp/'),
        data_format='csv',
        csv_files=['unlab.csv', 'amp_lab.csv', 'tox_lab.csv', 'sol_lab.csv', 'anticancer.csv',
                   'antihypertensive.csv', 'cell-cell.csv'],
        iteratorspecs=Bunch(
            train_vae=_weighted(['split=train']),
            train_amp_lab=_weighted(['split=train', 'amp']),
            hld_vae=_weighted(['split=val']),
            hld_unl=Bunch(subset=['split=val', '^amp']),
            hld_amppos=Bunch(subset=['split=val', 'amp=amp_posc,amp_posnc']),
            hld_ampneg=Bunch(subset=['split=val', 'amp=amp_negc,amp_negnc']),
        ),
        fixed_vocab_path=DATA_ROOT + 'amp/vocab.dict',
        split_seed=1288,
    ),
    data_prefixes=Bunch(dataset_type='bio', dataset_unl='amp_unlabeled', dataset_lab='amp_labeled'),
    attributes=[
        ('amp', {'amp_negnc': 0, 'amp_negc': 0, 'amp_posc': 1, 'amp_posnc': 1, 'na': -1}),
        ('tox', {'tox_negc': 0, 'tox_negnc': 0, 'tox_posc': 1, 'tox_posnc': 1, 'na': -1}),
        ('sol', {'sol_neg': 0, 'sol_pos': 1, 'na': -1}),
        ('anticancer', {'anticancer': 1, 'na': -1}),
        ('antihyper', {'antihyper': 1, 'na': -1}),
        ('hormone', {'cell': 1, 'na': -1}),
    ],
)


def _set_dataset(name):
    """Select the dataset spec (only 'amp' ships with the reference; 'yelp' is referenced but undefined)."""
    global data_kwargs, data_prefixes, attributes
    specs = {'amp': amp}
    if name not in specs:
        raise ValueError('unknown dataset ' + name)
    spec = specs[name]
    data_kwargs, data_prefixes, attributes = spec.data_kwargs, spec.data_prefixes, spec.attributes


def _update_cfg():
    """Post-process special values (paths, --tiny, part/partN, seeds, result file names)."""
    global savepath, tbpath, loadpath, vocab_path, seed, resume_result_json
    savepath = os.path.join(savepath_toplevel, runname)
    tbpath = os.path.join(tb_toplevel, runname)
    if tiny:                                     # fast smoke configuration (reference cfg.py:85-92)
        shared.update(n_iter=100, cheaplog_every=10, expsvlog_every=25, batch_size=5)
        evals.sample_size = 30
        full.s_iter = shared.n_iter
        resume_result_json = False
    if partN > 1:
        assert phase > 0, 'split in parts only makes sense when doing per-phase split'
        cfgv = vae if phase == 1 else full
        cfgv.n_iter = cfgv.n_iter // partN
        cfgv.s_iter += part * cfgv.n_iter
        cfgv.expsvlog_every = min(cfgv.expsvlog_every, cfgv.n_iter)
        assert (cfgv.s_iter + cfgv.n_iter) % cfgv.expsvlog_every == 0, \
            'Final model wont be saved; n_iter={}, expsvlog_every {}'.format(cfgv.n_iter, cfgv.expsvlog_every)
    vae.update(shared)
    full.update(shared)
    if vocab_path == 'auto':
        vocab_path = os.path.join(savepath, 'vocab.dict')
    chkpt = os.path.join(savepath, 'model_{}.pt')
    vae.chkpt_path = full.chkpt_path = chkpt
    if loadpath == 'auto':
        if part == 0 and phase != 2:
            loadpath = ''
        else:
            loadpath = chkpt.format((vae if phase == 1 else full).s_iter)
    if seed and phase > 0:
        seed += (phase - 1) * partN + part
    for bunch, files in ((vae, {'gen_samples_path': 'vae_gen.txt', 'eval_path': 'vae_eval.txt',
                                'fasta_gen_samples_path': 'vae_gen.fasta'}),
                         (full, {'gen_samples_path': 'full_gen.txt', 'samez_samples_path': 'full_samez.txt',
                                 'posz_samples_path': 'full_posz.txt', 'interp_samples_path': 'full_interp.txt',
                                 'eval_path': 'full_eval.txt', 'pos_eval_path': 'full.pos_eval.txt',
                                 'fasta_gen_samples_path': 'full_gen.fasta', 'fasta_pos_samples_path': 'pos_gen.fasta'})):
        for field, fn in files.items():
            bunch[field] = os.path.join(savepath, fn)
    _set_dataset(dataset)


_set_dataset(dataset)
