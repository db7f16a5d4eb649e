"""Decoder modules with the reference's names (models/decoder.py): GRUDecoder is the parameter
container of the 252->102 GRU and the Dropout+Linear(102->V) head (state_dict keys `rnn.*`,
`fc.1.*`, shared `emb`); WordDropout draws the <unk> mask exactly as the reference does
(numpy global RNG, every position eligible).  Arithmetic: csrc/gru.cu, dec_out.cu, decode.cu."""
import numpy as np
import torch
import torch.nn as nn

from models.mutils import UNK_IDX


def build_decoder(G_class, GRU_args, deconv_args, **common_args):
    if G_class == 'gru':
        kw = dict(GRU_args)
        kw.update(common_args)
        return GRUDecoder(**kw)
    if G_class == 'deconv':
        raise NotImplementedError("G_class='deconv' (DeconvDecoder) is outside the B200 hot path; use 'gru' "
                                  '(the reference default, cfg.py:276)')
    raise ValueError('Please use one of the following for dec_type: gru | deconv.')


class WordDropout(nn.Module):
    def __init__(self, p_word_dropout):
        super().__init__()
        self.p = p_word_dropout

    def sample_mask(self, shape):
        """uint8 mask, 1 = replace by <unk>; np.random.binomial like the reference (decoder.py:124-127)."""
        return torch.from_numpy(np.random.binomial(1, p=self.p, size=tuple(shape)).astype('uint8'))

    def forward(self, x):
        data = x.clone().detach()
        data[self.sample_mask(data.size()).to(x.device).bool()] = UNK_IDX
        return data


class GRUDecoder(nn.Module):
    def __init__(self, embedding, emb_dim, output_dim, h_dim, p_word_dropout, p_out_dropout, skip_connetions):
        super().__init__()
        if skip_connetions:
            raise NotImplementedError('skip_connetions=True is outside the B200 hot path (reference default False)')
        if not (emb_dim == 252 and h_dim == 102):
            raise NotImplementedError('cpg_b200 kernels are built for decoder input 252 / hidden 102')
        self.emb = embedding
        self.rnn = nn.GRU(emb_dim, h_dim, batch_first=True)
        self.fc = nn.Sequential(nn.Dropout(p_out_dropout), nn.Linear(h_dim, output_dim))
        self.word_dropout = WordDropout(p_word_dropout)
        self.skip_connetions = skip_connetions
        self.p_out_dropout = p_out_dropout

    def init_hidden(self, z, c):
        return torch.cat([z, c], dim=1)

    def forward(self, x, z, c):
        raise RuntimeError('GRUDecoder is evaluated through RNN_VAE.forward_decoder (fused kernels)')
