"""RNN_VAE with the reference's public surface (models/model.py of
IBM/controlled-peptide-generation), evaluated by the sm_100a kernels of libcpg_b200.

The module owns the same nn.Parameter containers as the reference, constructed in the same order
(word_emb, encoder, decoder, classifier -> identical initial weights under the same seed and an
identical state_dict key set), but the parameters are *views* into one flat fp32 buffer in the
layout the C ABI uses, so the fused training step (train_vae.py) and the module-level autograd
path below see the same storage.

What runs where
  forward()                 autograd.Function over cpg_wae_forward / cpg_wae_backward
  forward_encoder()         cpg_wae_encode (inference)
  forward_decoder()         cpg_wae_decode_teacher (inference)
  forward_classifier()      cpg_cnn_classifier_fwd (inference, eval-mode dropout)
  sample_G / generate_sentences   cpg_beam_decode / cpg_sample_decode
  flow_model (flow > 0)     cpg_flow_forward (models/flow.py), applied in generate_sentences like the reference
  soft sampling modes       cpg_soft_decode (none_softmax | greedy_softmax | categorical_softmax), forward only
Not built (raise NotImplementedError): soft (3-D) encoder / classifier inputs, prevent_empty, the deconv decoder, the
gumbel_* modes (placeholder strings in the reference as well) -- phase-2 training features of the reference.
"""
from itertools import chain

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from cpg_b200 import engine, sampling
from models.classifier import build_classifier
from models.decoder import build_decoder
from models.encoder import build_encoder
from models.mutils import EOS_IDX, PAD_IDX, START_IDX, UNK_IDX  # noqa: F401

_SOFT_MODES = ('gumbel_soft', 'gumbel_ST', 'greedy_softmax', 'categorical_softmax', 'none_softmax')


class _WaeFunction(torch.autograd.Function):
    """(tokens, noise, params...) -> (mu, logvar, z, logits); backward = BPTT kernels."""

    @staticmethod
    def forward(ctx, model, tokens, eps, c, word_drop, out_keep, p_out, *params):
        flat = model._flat
        need_grad = any(ctx.needs_input_grad)        # (grad mode is off inside Function.forward)
        mu, logvar, z, logits = engine.wae_forward(flat.params, flat.n_vocab, tokens, eps, c, word_drop, out_keep,
                                                   p_out, want_logits=True, keep_for_backward=need_grad)
        ctx.model, ctx.p_out = model, p_out
        ctx.stash_gen = engine.stash_generation(tokens.device) if need_grad else None
        ctx.save_for_backward(tokens, eps, c, word_drop, out_keep)
        ctx.n_params = len(params)
        return mu, logvar, z, logits

    @staticmethod
    def backward(ctx, d_mu, d_logvar, d_z, d_logits):
        tokens, eps, c, word_drop, out_keep = ctx.saved_tensors
        flat = ctx.model._flat
        cont = lambda t: None if t is None else t.contiguous()
        g = engine.wae_backward(flat.params, flat.n_vocab, tokens, eps, c, word_drop, out_keep, ctx.p_out,
                                cont(d_mu), cont(d_logvar), cont(d_z), cont(d_logits),
                                expect_generation=ctx.stash_gen)
        views = flat.views(g)
        return (None,) * 7 + tuple(views[name] for name in engine.PARAM_NAMES)


class RNN_VAE(nn.Module):
    def __init__(self, n_vocab, max_seq_len, z_dim, c_dim, emb_dim, pretrained_emb, freeze_embeddings, flow,
                 flow_type, E_args, G_args, C_args):
        super().__init__()
        if not (z_dim == 100 and c_dim == 2 and emb_dim == 150):
            raise NotImplementedError('cpg_b200 kernels are built for z_dim=100, c_dim=2, emb_dim=150')
        if not 4 <= n_vocab <= 32:
            raise NotImplementedError('cpg_b200 kernels support 4 <= n_vocab <= 32 (one warp lane per class)')
        self.MAX_SEQ_LEN = max_seq_len
        self.n_vocab, self.z_dim, self.c_dim, self.emb_dim = n_vocab, z_dim, c_dim, emb_dim
        self.device = torch.device('cuda')                       # as the reference (models/model.py:41)
        self.word_emb = nn.Embedding(n_vocab, emb_dim, PAD_IDX)
        if pretrained_emb is not None:
            assert emb_dim == pretrained_emb.size(1), 'emb dim dont match with pretrained'
            self.word_emb = nn.Embedding(n_vocab, emb_dim, PAD_IDX)
            self.word_emb.weight.data.copy_(pretrained_emb)
        if freeze_embeddings:
            self.word_emb.weight.requires_grad = False
        self.encoder = build_encoder('gru', emb_dim=emb_dim, z_dim=z_dim, **E_args)
        self.decoder = build_decoder(embedding=self.word_emb, emb_dim=emb_dim + z_dim + c_dim, output_dim=n_vocab,
                                     h_dim=z_dim + c_dim, **G_args)
        self.classifier = build_classifier('cnn', emb_dim, **C_args)
        self.use_flow = flow > 0                                  # reference models/model.py:70-73
        if self.use_flow:
            from models.flow import build_flow
            self.flow_model = build_flow(flow_type, flow, z_dim)
        self._flat = None
        self._decode_seed = 0

    # ------------------------------------------------------------------ flat parameter storage
    def _named_vae_params(self):
        own = dict(self.named_parameters())                       # decoder.emb.weight is de-duplicated away
        return [own[name] for name in engine.PARAM_NAMES]

    def flat_state(self):
        """FlatState whose `params` buffer the nn.Parameters alias (rebuilt after .to()/.cuda())."""
        ps = self._named_vae_params()
        dev = ps[0].device
        if not engine._lib._on_device(ps[0]):
            raise engine._lib.CpgLibraryError('RNN_VAE parameters are on %s; move the model to a CUDA device '
                                              '(there is no CPU path)' % dev)
        st = self._flat
        ok = st is not None and st.device == dev
        if ok:
            views = st.views(st.params)
            ok = all(p.data_ptr() == views[n].data_ptr() for p, n in zip(ps, engine.PARAM_NAMES))
        if not ok:
            old = st
            st = engine.FlatState(self.n_vocab, dev)
            views = st.views(st.params)
            with torch.no_grad():
                for p, name in zip(ps, engine.PARAM_NAMES):
                    views[name].copy_(p.data)
                    p.data = views[name]
            if old is not None and old.device == dev:           # keep optimizer moments across a rebuild
                st.adam_m.copy_(old.adam_m)
                st.adam_v.copy_(old.adam_v)
                st.step = old.step
            self._flat = st
        return st

    def bind_grads(self):
        """Point every VAE parameter's .grad at its slice of the flat gradient buffer."""
        st = self.flat_state()
        gv = st.views(st.grads)
        for p, name in zip(self._named_vae_params(), engine.PARAM_NAMES):
            p.grad = gv[name]
        return st

    # ------------------------------------------------------------------ parameter groups (model.py:75-94)
    def classifier_params(self):
        return filter(lambda p: p.requires_grad, self.classifier.parameters())

    def decoder_params(self):
        return filter(lambda p: p.requires_grad, self.decoder.parameters())

    def encoder_params(self):
        params = [self.word_emb.parameters(), self.encoder.parameters()]
        if self.use_flow:
            params.append(self.flow_model.parameters())
        return filter(lambda p: p.requires_grad, chain(*params))

    def vae_params(self):
        # word_emb is yielded twice (decoder.emb is word_emb), exactly like the reference
        params = [self.word_emb.parameters(), self.encoder.parameters(), self.decoder.parameters()]
        if self.use_flow:
            params.append(self.flow_model.parameters())
        return filter(lambda p: p.requires_grad, chain(*params))

    # ------------------------------------------------------------------ noise (reference RNG call sites)
    def sample_z(self, mu, logvar):
        eps = torch.randn(mu.size(0), self.z_dim).to(mu.device)
        return mu + torch.exp(logvar / 2) * eps

    def sample_z_prior(self, mbsize):
        return torch.randn(mbsize, self.z_dim).to(self._param_device())

    def sample_c_prior(self, mbsize):
        c = np.random.multinomial(1, [0.5, 0.5], mbsize).astype('float32')
        return torch.from_numpy(c).to(self._param_device())

    def _param_device(self):
        return self.word_emb.weight.device

    # ------------------------------------------------------------------ forwards
    def forward_encoder(self, inputs):
        if inputs.dim() != 2:
            raise NotImplementedError('soft (3-D) encoder inputs belong to phase 2 and are not built here')
        st = self.flat_state()
        return engine.wae_encode(st.params, self.n_vocab, inputs.contiguous())

    def forward_decoder(self, inputs, z, c):
        st = self.flat_state()
        B, L = inputs.shape
        wd = self.decoder.word_dropout.sample_mask((B, L)).to(inputs.device)
        keep = self._out_keep(B, L, inputs.device)
        logits = torch.empty(B, L, self.n_vocab, device=inputs.device)
        inp = engine._inputs(inputs.contiguous(), None, c.contiguous().float(), wd, keep, self.decoder.p_out_dropout)
        engine.check(engine.lib().cpg_wae_decode_teacher(engine.context(inputs.device), engine.stream_ptr(),
                                                         engine.ptr(st.params), self.n_vocab, B, L,
                                                         engine.byref(inp), engine.ptr(z.contiguous().float()),
                                                         engine.ptr(logits)), 'cpg_wae_decode_teacher')
        return logits

    def forward_classifier(self, inputs):
        if inputs.dim() != 2:
            raise NotImplementedError('soft (3-D) classifier inputs belong to phase 2 and are not built here')
        self.flat_state()
        cl = self.classifier
        return sampling.cnn_classifier_forward(self.word_emb.weight.data, [m.weight.data for m in cl.conv_layers],
                                               [m.bias.data for m in cl.conv_layers], cl.fc[1].weight.data,
                                               cl.fc[1].bias.data, inputs.contiguous())

    def _out_keep(self, B, L, device):
        p = self.decoder.p_out_dropout
        if not self.decoder.fc[0].training or p <= 0:
            return None
        return (torch.rand(B, L, self.z_dim + self.c_dim, device=device) >= p).to(torch.uint8)

    def forward(self, sequences, q_c='prior', sample_z=1):
        """-> ((mu, logvar), (z, c), dec_logits), reference models/model.py:146-195."""
        if self.use_flow:
            raise ValueError('Hmm if we properly want to do this, we need to compute flow and return flow-kl loss out of '
                             'function.')                            # as the reference, models/model.py:173-176
        mbsize, L = sequences.shape
        dev = sequences.device
        st = self.flat_state()
        if sample_z == 'max':
            eps = None
        else:
            assert sample_z == 1, 'sample_z > 1 is a TODO in the reference as well'
            eps = torch.randn(mbsize, self.z_dim).to(dev)          # CPU generator, like model.py:111
        if isinstance(q_c, torch.Tensor):
            c = torch.zeros(mbsize, 2, device=dev).scatter_(1, q_c.unsqueeze(1), 1)
        elif q_c == 'prior':
            c = self.sample_c_prior(mbsize)
        elif q_c == 'classifier':
            c = F.softmax(self.forward_classifier(sequences), dim=1)
        else:
            raise ValueError('q_c is not labels, prior, or classifier')
        wd = self.decoder.word_dropout.sample_mask((mbsize, L)).to(dev)     # drawn in eval mode too (decoder.py:117)
        keep = self._out_keep(mbsize, L, dev)
        mu, logvar, z, logits = _WaeFunction.apply(self, sequences.contiguous(), eps, c.contiguous(), wd, keep,
                                                   self.decoder.p_out_dropout, *self._named_vae_params())
        assert st is self._flat
        return (mu, logvar), (z, c), logits

    # ------------------------------------------------------------------ generation (model.py:197-385)
    def generate_sentences(self, mbsize, z=None, c=None, eval_mode=True, **sample_kwargs):
        if z is None:
            z = self.sample_z_prior(mbsize)
        if c is None:
            c = self.sample_c_prior(mbsize)
        if self.use_flow:                                         # reference :210-214 (train flag = eval_mode)
            z = z.to(self._param_device())
            z = self.flow_model(z, train=eval_mode)
            if eval_mode:
                z = z[0]
        sentences = self.sample_G(mbsize, z, c, **sample_kwargs)
        return sentences, z, c.argmax(dim=1)

    def sample_G(self, mbsize, z, c, sample_mode='categorical', temp=1.0, gumbel_temp=1.0, prepend_start_idx=True,
                 prevent_empty=False, min_length=1, beam_size=5, n_best=3):
        if sample_mode in ('gumbel_soft', 'gumbel_ST', 'gumbel_max'):
            raise NotImplementedError("sample_mode '%s' is a placeholder string in the reference too (models/model.py:313,"
                                      '331-335); implemented: categorical | greedy | beam | none_softmax | greedy_softmax | '
                                      'categorical_softmax' % sample_mode)
        assert not (sample_mode in _SOFT_MODES and prevent_empty), 'cant prevent_empty when soft sampling'
        if prevent_empty or min_length != 1:
            raise NotImplementedError('prevent_empty / min_length != 1 are not built into the decode kernels')
        assert beam_size >= n_best, "Can't return more than max hypothesis"
        assert mbsize == z.size(0) == c.size(0), 'oops sizes dont match {} {} {}'.format(mbsize, z.size(0), c.size(0))
        st = self.flat_state()
        dev = st.device
        z, c = z.to(dev).float(), c.to(dev).float()
        L = self.MAX_SEQ_LEN
        if sample_mode == 'beam':
            toks, lens, _ = sampling.beam_decode(st.params, self.n_vocab, z, c, L, beam_size, n_best)
            toks, lens = toks.cpu().tolist(), lens.cpu().tolist()
            return [[toks[j][i][:lens[j][i]] for i in range(n_best)] for j in range(mbsize)]
        if sample_mode in sampling.SOFT_MODES:
            # forward only: the soft samples are not differentiated (phase-2 training is outside this package)
            seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if sample_mode == 'categorical_softmax' else 0
            seq, soft = sampling.soft_decode(st.params, self.n_vocab, z, c, sample_mode, L, temp, seed)
            return (seq, soft) if prepend_start_idx else (seq[:, 1:], soft[:, 1:])
        if sample_mode == 'greedy':
            seq = sampling.sample_decode(st.params, self.n_vocab, z, c, sampling.MODE_GREEDY, L)
        elif sample_mode == 'categorical':
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())       # tie the Philox key to torch's seeded stream
            seq = sampling.sample_decode(st.params, self.n_vocab, z, c, sampling.MODE_CATEGORICAL, L, temp, seed)
        else:
            raise Exception('Sample mode {} not implemented.'.format(sample_mode))
        return seq if prepend_start_idx else seq[:, 1:]
