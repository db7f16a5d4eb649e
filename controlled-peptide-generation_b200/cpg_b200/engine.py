"""Thin functional layer over the C ABI: flat parameter state, the WAE forward /
backward / fused-step calls and the individual loss ops.  All tensors are torch
CUDA tensors owned by the caller; this module only marshals pointers.
"""
import contextlib
import math
from collections import OrderedDict
from ctypes import byref, c_void_p

import torch

from . import _lib
from ._lib import (SC, SC_COUNT, ZREGU, LossNoise, StepNoiseBuffers, TrainHparams, WaeInputs, check, context, lib, ptr,
                   stream_ptr)

# layout order of cpg_vae_param_layout (== unique tensors of RNN_VAE.vae_params(), models/model.py:88-94)
PARAM_NAMES = (
    'word_emb.weight',
    'encoder.rnn.weight_ih_l0', 'encoder.rnn.weight_hh_l0', 'encoder.rnn.bias_ih_l0', 'encoder.rnn.bias_hh_l0',
    'encoder.rnn.weight_ih_l0_reverse', 'encoder.rnn.weight_hh_l0_reverse',
    'encoder.rnn.bias_ih_l0_reverse', 'encoder.rnn.bias_hh_l0_reverse',
    'encoder.q_mu.weight', 'encoder.q_mu.bias', 'encoder.q_logvar.weight', 'encoder.q_logvar.bias',
    'decoder.rnn.weight_ih_l0', 'decoder.rnn.weight_hh_l0', 'decoder.rnn.bias_ih_l0', 'decoder.rnn.bias_hh_l0',
    'decoder.fc.1.weight', 'decoder.fc.1.bias',
)
EMB, ENC_H, ZD, CD, DEC_H = 150, 80, 100, 2, 102
DP_TAIL = 8          # == cpg_dp_tail_count()
STATIC_STREAM = 0x80000000


def param_shapes(n_vocab):
    ge, gd = 3 * ENC_H, 3 * DEC_H
    shapes = [(n_vocab, EMB),
              (ge, EMB), (ge, ENC_H), (ge,), (ge,), (ge, EMB), (ge, ENC_H), (ge,), (ge,),
              (ZD, 2 * ENC_H), (ZD,), (ZD, 2 * ENC_H), (ZD,),
              (gd, EMB + DEC_H), (gd, DEC_H), (gd,), (gd,),
              (n_vocab, DEC_H), (n_vocab,)]
    return OrderedDict(zip(PARAM_NAMES, shapes))


class FlatState:
    """Flat fp32 buffers (params, grads, Adam m/v) in the library's layout, with
    named views so that nn.Parameters / state_dicts can alias them."""

    def __init__(self, n_vocab, device=None):
        self.n_vocab = int(n_vocab)
        self.device = _lib.tensor_device(device)
        self.offsets, self.sizes, self.total = _lib.param_layout(self.n_vocab)
        self.shapes = param_shapes(self.n_vocab)
        z = lambda: torch.zeros(self.total, dtype=torch.float32, device=self.device)
        self.params, self.adam_m, self.adam_v = z(), z(), z()
        # the flat gradient is followed by a few floats that ride along in the data-parallel gradient all-reduce
        # (parallel.py): one buffer, `grads` is the view the kernels and nn.Parameter.grad see
        self.grads_ext = torch.zeros(self.total + DP_TAIL, dtype=torch.float32, device=self.device)
        self.grads = self.grads_ext[:self.total]
        self.step = 0

    def reset_optimizer(self):
        """Fresh Adam state, as a newly constructed optim.Adam (train_vae.py:15 builds one per call)."""
        self.adam_m.zero_()
        self.adam_v.zero_()
        self.step = 0

    def views(self, flat):
        out = OrderedDict()
        for name, off, n in zip(PARAM_NAMES, self.offsets, self.sizes):
            out[name] = flat[off:off + n].view(self.shapes[name])
        return out

    def load(self, named_tensors):
        """Copy tensors (dict keyed by the reference's state_dict names) into the flat params."""
        v = self.views(self.params)
        with torch.no_grad():
            for name in PARAM_NAMES:
                src = named_tensors[name] if name in named_tensors else named_tensors['decoder.emb.weight']
                v[name].copy_(src.to(self.device, torch.float32))

    def export(self, flat=None):
        return OrderedDict((k, t.clone()) for k, t in self.views(self.params if flat is None else flat).items())


def _inputs(tokens, eps, c, word_drop, out_keep, p_out):
    inp = WaeInputs()
    inp.tokens = ptr(tokens, torch.int64, allow_none=False)
    inp.eps = ptr(eps, torch.float32)
    inp.c = ptr(c, torch.float32)
    inp.word_drop = ptr(word_drop, torch.uint8)
    inp.out_keep = ptr(out_keep, torch.uint8)
    inp.p_out_dropout = float(p_out)
    return inp


def _check_shapes(tokens, eps, c, word_drop, out_keep):
    if tokens.dim() != 2:
        raise ValueError('tokens must be [B, L]')
    B, L = tokens.shape
    for name, t, shp in (('eps', eps, (B, ZD)), ('c', c, (B, CD)), ('word_drop', word_drop, (B, L)),
                         ('out_keep', out_keep, (B, L, DEC_H))):
        if t is not None and tuple(t.shape) != shp:
            raise ValueError('%s must have shape %s, got %s' % (name, shp, tuple(t.shape)))
    return B, L


def wae_forward(params, n_vocab, tokens, eps, c, word_drop=None, out_keep=None, p_out=0.3, want_logits=True,
                keep_for_backward=False):
    """RNN_VAE.forward arithmetic (models/model.py:146-195) on explicit noise."""
    B, L = _check_shapes(tokens, eps, c, word_drop, out_keep)
    dev = tokens.device
    mu = torch.empty(B, ZD, device=dev)
    logvar = torch.empty(B, ZD, device=dev)
    z = torch.empty(B, ZD, device=dev)
    logits = torch.empty(B, L, n_vocab, device=dev) if want_logits else None
    inp = _inputs(tokens, eps, c, word_drop, out_keep, p_out)
    check(lib().cpg_wae_forward(context(dev), stream_ptr(), ptr(params), n_vocab, B, L, byref(inp), ptr(mu),
                                ptr(logvar), ptr(z), ptr(logits), 1 if keep_for_backward else 0), 'cpg_wae_forward')
    return mu, logvar, z, logits


def stash_generation(device):
    """Id of the BPTT stash the context holds right now (-1: none); see cpg_stash_generation."""
    return int(lib().cpg_stash_generation(context(device)))


def wae_backward(params, n_vocab, tokens, eps, c, word_drop, out_keep, p_out, d_mu, d_logvar, d_z, d_logits,
                 grads_out=None, expect_generation=None):
    B, L = _check_shapes(tokens, eps, c, word_drop, out_keep)
    dev = tokens.device
    if expect_generation is not None and stash_generation(dev) != expect_generation:
        raise _lib.CpgLibraryError(
            'backward of a forward whose activation stash is gone: another forward / decode / encode ran on this '
            'device in between (the context keeps ONE stash; run forward -> backward pairs one at a time)')
    if grads_out is None:
        grads_out = torch.empty_like(params)
    inp = _inputs(tokens, eps, c, word_drop, out_keep, p_out)
    check(lib().cpg_wae_backward(context(dev), stream_ptr(), ptr(params), n_vocab, B, L, byref(inp), ptr(d_mu),
                                 ptr(d_logvar), ptr(d_z), ptr(d_logits), ptr(grads_out)), 'cpg_wae_backward')
    return grads_out


def wae_encode(params, n_vocab, tokens):
    B, L = tokens.shape
    dev = tokens.device
    mu = torch.empty(B, ZD, device=dev)
    logvar = torch.empty(B, ZD, device=dev)
    check(lib().cpg_wae_encode(context(dev), stream_ptr(), ptr(params), n_vocab, B, L,
                               ptr(tokens, torch.int64, allow_none=False), ptr(mu), ptr(logvar)), 'cpg_wae_encode')
    return mu, logvar


def make_hparams(lr=1e-3, betas=(0.9, 0.999), adam_eps=1e-8, clip_norm=5.0, beta=1.0, lambda_logvar_l1=0.0,
                 lambda_logvar_kl=1e-3, z_regu='mmdrf', mmd_sigma=7.0, rf_dim=500, compute_full_mmd=True,
                 adam_step=1, global_batch=0):
    hp = TrainHparams()
    hp.lr, hp.beta1, hp.beta2, hp.adam_eps = lr, betas[0], betas[1], adam_eps
    hp.clip_norm, hp.beta = clip_norm, beta
    hp.lambda_logvar_l1, hp.lambda_logvar_kl = lambda_logvar_l1, lambda_logvar_kl
    hp.z_regu = ZREGU[z_regu] if isinstance(z_regu, str) else int(z_regu)
    hp.mmd_sigma, hp.rf_dim = mmd_sigma, rf_dim
    hp.compute_full_mmd = 1 if compute_full_mmd else 0
    hp.adam_step, hp.global_batch = adam_step, global_batch
    return hp


def _loss_noise(noise):
    nz = LossNoise()
    nz.z_prior_full = ptr(noise.get('z_prior_full'), torch.float32)
    nz.z_prior_rf = ptr(noise['z_prior_rf'], torch.float32, allow_none=False)
    nz.rf_w = ptr(noise['rf_w'], torch.float32, allow_none=False)
    nz.rf_b = ptr(noise['rf_b'], torch.float32, allow_none=False)
    return nz


def train_step(state, tokens, noise, hp, p_out=0.3, want=()):
    """One fused iteration of train_vae.py:24-42 on one GPU.  `noise` holds eps, c,
    word_drop, out_keep, z_prior_full, z_prior_rf, rf_w, rf_b (device tensors).
    Returns (scalars[SC_COUNT] device tensor, dict of requested extras)."""
    B, L = _check_shapes(tokens, noise['eps'], noise['c'], noise.get('word_drop'), noise.get('out_keep'))
    dev = tokens.device
    V = state.n_vocab
    scalars = torch.zeros(SC_COUNT, device=dev)
    extras = {}
    for k, shp in (('mu', (B, ZD)), ('logvar', (B, ZD)), ('z', (B, ZD)), ('logits', (B, L, V))):
        extras[k] = torch.empty(shp, device=dev) if k in want else None
    inp = _inputs(tokens, noise['eps'], noise['c'], noise.get('word_drop'), noise.get('out_keep'), p_out)
    nz = _loss_noise(noise)
    state.step += 1
    hp.adam_step = state.step
    check(lib().cpg_wae_train_step(context(dev), stream_ptr(), ptr(state.params), ptr(state.grads), ptr(state.adam_m),
                                   ptr(state.adam_v), V, B, L, byref(inp), byref(nz), byref(hp), ptr(scalars),
                                   ptr(extras['mu']), ptr(extras['logvar']), ptr(extras['z']), ptr(extras['logits'])),
          'cpg_wae_train_step')
    return scalars, {k: v for k, v in extras.items() if v is not None}


def step_phase1(state, tokens, noise, hp, p_out=0.3):
    B, L = _check_shapes(tokens, noise['eps'], noise['c'], noise.get('word_drop'), noise.get('out_keep'))
    dev = tokens.device
    coupled = torch.zeros(int(lib().cpg_coupled_count(hp.rf_dim)), device=dev)
    inp = _inputs(tokens, noise['eps'], noise['c'], noise.get('word_drop'), noise.get('out_keep'), p_out)
    nz = _loss_noise(noise)
    z = torch.empty(B, ZD, device=dev)
    check(lib().cpg_wae_step_phase1(context(dev), stream_ptr(), ptr(state.params), state.n_vocab, B, L, byref(inp),
                                    byref(nz), byref(hp), ptr(coupled), c_void_p(None), c_void_p(None), ptr(z)),
          'cpg_wae_step_phase1')
    return coupled, z


def step_phase2(state, tokens, noise, hp, coupled, p_out=0.3):
    B, L = tokens.shape
    dev = tokens.device
    scalars = torch.zeros(SC_COUNT, device=dev)
    inp = _inputs(tokens, noise['eps'], noise['c'], noise.get('word_drop'), noise.get('out_keep'), p_out)
    nz = _loss_noise(noise)
    check(lib().cpg_wae_step_phase2(context(dev), stream_ptr(), ptr(state.params), ptr(state.grads), state.n_vocab,
                                    B, L, byref(inp), byref(nz), byref(hp), ptr(coupled), ptr(scalars),
                                    c_void_p(None)), 'cpg_wae_step_phase2')
    return scalars


def clip_adam(state, hp, out=None):
    """clip_grad_norm_ + Adam on the flat buffers; the pre-clip norm goes to `out` (1 float, device)."""
    dev = state.params.device
    gn = torch.zeros(1, device=dev) if out is None else out
    check(lib().cpg_clip_adam_step(context(dev), stream_ptr(), ptr(state.params), ptr(state.grads), ptr(state.adam_m),
                                   ptr(state.adam_v), state.n_vocab, byref(hp), ptr(gn)), 'cpg_clip_adam_step')
    return gn


# ------------------------------------------------------------------ data-parallel plumbing (parallel.py)
def dp_tail_count():
    return DP_TAIL


def side_stream(device):
    """The library's internal side stream as a torch stream (None when option side_stream = 0)."""
    p = lib().cpg_side_stream(context(device))
    return torch.cuda.ExternalStream(p, device=device) if p else None


def stats_stream(device):
    """Context manager: work enqueued inside is ordered behind the stream that produced phase 1's `coupled[1:]`."""
    s = side_stream(device)
    return torch.cuda.stream(s) if s is not None else contextlib.nullcontext()


def ntok_stream(device):
    """Context manager: work enqueued inside is ordered behind the stream that produced phase 1's `coupled[0]`
    (the token count, available right after the token preparation)."""
    p = lib().cpg_aux_stream(context(device))
    return torch.cuda.stream(torch.cuda.ExternalStream(p, device=device)) if p else contextlib.nullcontext()


def join_coupled(device):
    """Make the current stream wait for the library's internal streams: a caller that reads phase 1's `coupled` on its own
    stream (instead of exchanging it on stats_stream / ntok_stream) calls this first."""
    cur = torch.cuda.current_stream(device)
    for getter in (lib().cpg_side_stream, lib().cpg_aux_stream):
        p = getter(context(device))
        if p:
            cur.wait_stream(torch.cuda.ExternalStream(p, device=device))


def step_dyn_write(device, hp, noise_step):
    """Refresh the per-step scalars (beta, Adam bias corrections of hp.adam_step, noise counter) in the context's device
    block: what the kernels of a caller-captured iteration read (see step_dyn_use)."""
    check(lib().cpg_step_dyn_write(context(device), stream_ptr(), byref(hp), int(noise_step)), 'cpg_step_dyn_write')


def step_dyn_use(device, on):
    check(lib().cpg_step_dyn_use(context(device), 1 if on else 0), 'cpg_step_dyn_use')


def dp_pack_tail(tail):
    check(lib().cpg_dp_pack_tail(context(tail.device), stream_ptr(), ptr(tail)), 'cpg_dp_pack_tail')


def dp_apply_tail(tail, scalars):
    check(lib().cpg_dp_apply_tail(context(tail.device), stream_ptr(), ptr(tail), ptr(scalars)), 'cpg_dp_apply_tail')


# ------------------------------------------------------------------ losses.py ops
def softmax_xent(logits, tokens, want_grad=False):
    B, L, V = logits.shape
    dev = logits.device
    out = torch.empty(2, device=dev)
    dl = torch.empty_like(logits) if want_grad else None
    check(lib().cpg_softmax_xent(context(dev), stream_ptr(), ptr(logits.contiguous(), torch.float32),
                                 ptr(tokens, torch.int64), B, L, V, ptr(out), ptr(dl)), 'cpg_softmax_xent')
    return out, dl


def latent_stats(mu, logvar):
    out = torch.empty(5, device=mu.device)
    check(lib().cpg_latent_stats(context(mu.device), stream_ptr(), ptr(mu.contiguous(), torch.float32),
                                 ptr(logvar.contiguous(), torch.float32), mu.shape[0], ptr(out)), 'cpg_latent_stats')
    return out


def mmd_full(z, z_prior, sigma, out=None):
    out = torch.empty(1, device=z.device) if out is None else out
    check(lib().cpg_mmd_full(context(z.device), stream_ptr(), ptr(z.contiguous(), torch.float32),
                             ptr(z_prior.contiguous(), torch.float32), z.shape[0], float(sigma), ptr(out)),
          'cpg_mmd_full')
    return out


def mmd_full_grad(z, z_prior, sigma):
    dz = torch.empty_like(z, dtype=torch.float32)
    check(lib().cpg_mmd_full_grad(context(z.device), stream_ptr(), ptr(z.contiguous(), torch.float32),
                                  ptr(z_prior.contiguous(), torch.float32), z.shape[0], float(sigma), ptr(dz)), 'cpg_mmd_full_grad')
    return dz


def mmd_rf(z, z_prior, rf_w, rf_b, sigma, want_grad=False):
    out = torch.empty(1, device=z.device)
    dz = torch.empty_like(z) if want_grad else None
    check(lib().cpg_mmd_rf(context(z.device), stream_ptr(), ptr(z.contiguous(), torch.float32),
                           ptr(z_prior.contiguous(), torch.float32), ptr(rf_w.contiguous(), torch.float32),
                           ptr(rf_b.contiguous(), torch.float32), z.shape[0], rf_w.shape[1], float(sigma), ptr(out),
                           ptr(dz)), 'cpg_mmd_rf')
    return out, dz


# ------------------------------------------------------------------ perf-mode noise + trainer
def fill_step_noise(noise, seed, step, p_word=0.3, p_out=0.3, overlap=False):
    """Regenerate the per-iteration noise tensors in place (Philox; pure function of seed/step).
    overlap=True: the request is only RECORDED; the train-step / forward entry point that reads these buffers next draws the
    word-dropout mask inside its token preparation and the rest on the library's side stream -- only when the next reader IS
    such an entry point (do not read the tensors from Python in between: use overlap=False for that)."""
    B, L = noise['word_drop'].shape
    dev = noise['eps'].device
    fn = lib().cpg_fill_step_noise_overlapped if overlap else lib().cpg_fill_step_noise
    check(fn(context(dev), stream_ptr(), int(seed), int(step), B, L, float(p_word),
                                    float(p_out), ptr(noise['eps']), ptr(noise['c']), ptr(noise['word_drop']),
                                    ptr(noise['out_keep']), ptr(noise.get('z_prior_full')),
                                    ptr(noise['z_prior_rf'])), 'cpg_fill_step_noise')


def fill_normal(out, seed, stream_id):
    check(lib().cpg_fill_normal(context(out.device), stream_ptr(), int(seed), int(stream_id), out.numel(), ptr(out)),
          'cpg_fill_normal')
    return out


def fill_uniform(out, seed, stream_id, scale=1.0):
    check(lib().cpg_fill_uniform(context(out.device), stream_ptr(), int(seed), int(stream_id), float(scale),
                                 out.numel(), ptr(out)), 'cpg_fill_uniform')
    return out


def alloc_noise(B, L, device, rf_dim=500, seed=1238, full_mmd=True):
    """Device buffers for one iteration's noise; rf_w / rf_b are drawn once (losses.py:73-81 caches them)."""
    dev = _lib.tensor_device(device)
    noise = {
        'eps': torch.empty(B, ZD, device=dev),
        'c': torch.empty(B, CD, device=dev),
        'word_drop': torch.empty(B, L, dtype=torch.uint8, device=dev),
        'out_keep': torch.empty(B, L, DEC_H, dtype=torch.uint8, device=dev),
        'z_prior_rf': torch.empty(B, ZD, device=dev),
        'rf_w': torch.empty(ZD, rf_dim, device=dev),
        'rf_b': torch.empty(rf_dim, device=dev),
    }
    if full_mmd:
        noise['z_prior_full'] = torch.empty(B, ZD, device=dev)
    # static draws live in the upper half of the Philox stream-id space; the per-step noise uses 16*step + k < 2**31
    fill_normal(noise['rf_w'], seed, STATIC_STREAM | 0xF001)
    fill_uniform(noise['rf_b'], seed, STATIC_STREAM | 0xF002, scale=2 * math.pi)
    return noise


class FusedStepper:
    """Perf-mode iteration with everything pre-marshalled: the noise buffers, the scalar block and the
    ctypes argument structs are created once, so a step costs two C calls (Philox noise + fused
    iteration) and no allocation."""

    def __init__(self, state, B, L, hp, seed=1238, p_word=0.3, p_out=0.3, rf_dim=500):
        self.state, self.hp, self.seed = state, hp, int(seed)
        self.B, self.L, self.p_word, self.p_out = B, L, float(p_word), float(p_out)
        dev = state.device
        self.noise = alloc_noise(B, L, dev, rf_dim=rf_dim, seed=seed)
        # two scalar blocks used in turn (each has its captured graph): a reader may still be copying step k's block to the
        # host on another stream while step k + 1 composes its own
        self._scal = [torch.zeros(SC_COUNT, device=dev), torch.zeros(SC_COUNT, device=dev)]
        self._flip = 0
        self.scalars = self._scal[0]
        self.ctx, self.lib = context(dev), lib()
        self.nz = _loss_noise(self.noise)
        self._nb = None
        n = self.noise
        self._noise_args = (ptr(n['eps']), ptr(n['c']), ptr(n['word_drop']), ptr(n['out_keep']),
                            ptr(n.get('z_prior_full')), ptr(n['z_prior_rf']))
        self._bufs = (ptr(state.params), ptr(state.grads), ptr(state.adam_m), ptr(state.adam_v))
        self._scal_ptrs = [ptr(t) for t in self._scal]
        self._null = c_void_p(None)

    def step(self, tokens, it, beta):
        """tokens: contiguous int64 [B, L] on the device.  Returns the device scalar block (not synchronised).
        One C call: Philox noise + the fused iteration, replayed from a captured CUDA graph after two eager steps."""
        st, hp, n = self.state, self.hp, self.noise
        if self._nb is None:
            nb = StepNoiseBuffers()
            nb.eps, nb.c, nb.word_drop, nb.out_keep = ptr(n['eps']), ptr(n['c']), ptr(n['word_drop']), ptr(n['out_keep'])
            nb.z_prior_full, nb.z_prior_rf = ptr(n.get('z_prior_full')), ptr(n['z_prior_rf'])
            nb.rf_w, nb.rf_b = ptr(n['rf_w']), ptr(n['rf_b'])
            self._nb = nb
        st.step += 1
        hp.adam_step = st.step
        hp.beta = float(beta)
        p, g, m, v = self._bufs
        self.scalars, sc = self._scal[self._flip], self._scal_ptrs[self._flip]
        self._flip ^= 1
        check(self.lib.cpg_wae_train_step_philox(self.ctx, stream_ptr(), p, g, m, v, st.n_vocab, self.B, self.L,
                                                 ptr(tokens, torch.int64, allow_none=False), byref(self._nb), byref(hp), self.seed,
                                                 int(it), self.p_word, self.p_out, sc), 'cpg_wae_train_step_philox')
        return self.scalars
