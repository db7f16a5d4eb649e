"""Live-reference harness: imports the UNMODIFIED reference from /root/reference.

TEST INFRASTRUCTURE ONLY, and usable only in the build container (the GPU box
has no /root/reference).  It is used by `oracle/gen_golden.py` to produce the
committed fixtures under `tests/golden/`, and by `bench.py --impl reference`
when /root/reference happens to exist.  No reference source is copied: the
modules are imported from where they lie, with import stubs for the packages
the image lacks and with the third-party RNG entry
points (torch.randn, np.random.binomial, F.dropout ...) wrapped so that the
random tensors of an iteration are *injected* from explicit arrays.
"""
import contextlib
import math
import os
import sys
import types

import numpy as np
import torch

REF_ROOT = os.environ.get('CPG_REFERENCE_ROOT', '/root/reference')


def reference_available():
    return os.path.isfile(os.path.join(REF_ROOT, 'train_vae.py'))


class _LogSink:
    """Collects tb_json_logger.log_value(name, value, step) calls."""
    def __init__(self):
        self.values = {}

    def log_value(self, name, value, step=None):
        self.values.setdefault(step, {})[name] = float(value)


LOG = _LogSink()


def _install_stubs():
    tl = types.ModuleType('tensorboard_logger')
    tl2 = types.ModuleType('tensorboard_logger.tensorboard_logger')

    class Logger:
        def __init__(self, *a, **k):
            pass

        def log_value(self, *a, **k):
            pass
        log_histogram = log_images = log_value
    for n in ('configure', 'log_value', 'log_histogram', 'log_images'):
        setattr(tl2, n, lambda *a, **k: None)
        setattr(tl, n, lambda *a, **k: None)
    tl2.Logger = Logger
    tl.Logger = Logger
    tl.tensorboard_logger = tl2
    sys.modules.setdefault('tensorboard_logger', tl)
    sys.modules.setdefault('tensorboard_logger.tensorboard_logger', tl2)
    for m in ('h5py', 'matplotlib', 'matplotlib.pyplot', 'seaborn'):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules['matplotlib.pyplot'].rc = lambda *a, **k: None
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']


_REF = None


def load_reference():
    """Import the reference modules (once) and return them in a namespace."""
    global _REF
    if _REF is not None:
        return _REF
    if not reference_available():
        raise RuntimeError('reference tree not found at %s' % REF_ROOT)
    _install_stubs()
    saved_argv = sys.argv
    sys.argv = ['x']
    sys.path.insert(0, REF_ROOT)
    try:
        import cfg
        cfg._update_cfg()
        import losses
        import train_vae
        import density_modeling
        from models import model as model_mod
        from models import decoder as decoder_mod
        from models import classifier as classifier_mod
    finally:
        sys.argv = saved_argv
    train_vae.log_value = LOG.log_value          # where the scalars of train_vae.py:44-53 land
    _REF = types.SimpleNamespace(cfg=cfg, losses=losses, train_vae=train_vae,
                                 density_modeling=density_modeling, model=model_mod,
                                 decoder=decoder_mod, classifier=classifier_mod)
    return _REF


def build_model(n_vocab, seed=1238):
    """main.py:44-46,63-64 on CPU (device override as api.py:96 does)."""
    ref = load_reference()
    torch.manual_seed(seed)
    np.random.seed(seed)
    m = ref.model.RNN_VAE(n_vocab=n_vocab, max_seq_len=ref.cfg.max_seq_len, **ref.cfg.model)
    m.device = torch.device('cpu')
    return m


class DatasetShim:
    """Stands in for AttributeDataLoader (data_processing/dataset.py:285-300):
    `next_batch(name).text` is int64 [B, 25]; `idx2sentence(s)` joins ids."""
    def __init__(self, batches, on_batch=None):
        self.batches = list(batches)
        self.i = 0
        self.on_batch = on_batch

    def next_batch(self, name):
        if self.on_batch is not None:
            self.on_batch(self.i)
        b = self.batches[min(self.i, len(self.batches) - 1)]
        self.i += 1
        return types.SimpleNamespace(text=b)

    def idx2sentence(self, idxs, print_special_tokens=True):
        return ' '.join(str(int(i)) for i in idxs.view(-1))

    def idx2sentences(self, seqs, print_special_tokens=True):
        return [' '.join(str(int(i)) for i in s if print_special_tokens or int(i) > 3) for s in seqs]


class NoiseInjector:
    """Feeds pre-drawn noise into the reference's RNG call sites (appendix A).

    Queues (consumed in call order):
      randn[shape]      <- torch.randn(*shape)               (model.py:111,118; losses.py:75)
      randn_like[shape] <- torch.randn_like(x)               (losses.py:37; density_modeling.py:67)
      rand[shape]       <- torch.rand(shape)                 (losses.py:76)
      multinomial       <- np.random.multinomial(1,[.5,.5],B) (model.py:125)
      binomial          <- np.random.binomial(1,p,size)      (decoder.py:124-127)
      dropout           <- F.dropout keep masks              (decoder.py:43-45 / classifier.py:33)
    Calls whose queue is empty fall through to the real generator.
    """
    def __init__(self):
        self.q = {'randn': {}, 'randn_like': {}, 'rand': {}, 'multinomial': [], 'binomial': [],
                  'dropout': []}

    def push(self, kind, value):
        if kind in ('randn', 'randn_like', 'rand'):
            self.q[kind].setdefault(tuple(value.shape), []).append(value)
        else:
            self.q[kind].append(value)

    @contextlib.contextmanager
    def active(self):
        import torch.nn.functional as F
        real = dict(randn=torch.randn, randn_like=torch.randn_like, rand=torch.rand,
                    multinomial=np.random.multinomial, binomial=np.random.binomial,
                    dropout=F.dropout)
        q = self.q

        def _shape(args):
            if len(args) == 1 and isinstance(args[0], (tuple, list, torch.Size)):
                return tuple(args[0])
            return tuple(int(a) for a in args)

        def randn(*args, **kw):
            lst = q['randn'].get(_shape(args))
            if lst:
                return lst.pop(0).clone()
            return real['randn'](*args, **kw)

        def rand(*args, **kw):
            lst = q['rand'].get(_shape(args))
            if lst:
                return lst.pop(0).clone()
            return real['rand'](*args, **kw)

        def randn_like(x, **kw):
            lst = q['randn_like'].get(tuple(x.shape))
            if lst:
                return lst.pop(0).clone().to(x.dtype)
            return real['randn_like'](x, **kw)

        def multinomial(n, pvals, size=None):
            if q['multinomial']:
                return np.asarray(q['multinomial'].pop(0))
            return real['multinomial'](n, pvals, size)

        def binomial(n, p, size=None):
            if q['binomial']:
                return np.asarray(q['binomial'].pop(0))
            return real['binomial'](n, p, size)

        def dropout(inp, p=0.5, training=True, inplace=False):
            if not training:
                return inp
            if q['dropout']:
                keep = q['dropout'].pop(0).to(inp.dtype)
                return inp * keep * (1.0 / (1.0 - p))
            return real['dropout'](inp, p, training, inplace)

        torch.randn, torch.randn_like, torch.rand = randn, randn_like, rand
        np.random.multinomial, np.random.binomial = multinomial, binomial
        F.dropout = dropout
        try:
            yield self
        finally:
            torch.randn, torch.randn_like, torch.rand = real['randn'], real['randn_like'], real['rand']
            np.random.multinomial, np.random.binomial = real['multinomial'], real['binomial']
            F.dropout = real['dropout']


def push_iteration_noise(inj, noise, first_iteration):
    """Queue one training iteration's noise in the reference's draw order."""
    inj.push('randn', noise['eps'])
    inj.push('multinomial', noise['c'].numpy().astype('int64'))
    inj.push('binomial', noise['word_drop'].numpy())
    inj.push('dropout', noise['out_keep'])
    inj.push('randn_like', noise['z_prior_full'])
    inj.push('randn_like', noise['z_prior_rf'])
    if first_iteration:
        inj.push('randn', noise['rf_w'])
        inj.push('rand', noise['rf_u'])


def reset_rf_cache():
    """losses.py:66 keeps (rf_w, rf_b) in a module global; clear it between runs."""
    load_reference().losses.rf.clear()
