"""Encoder module with the reference's class / attribute names (models/encoder.py): a parameter
container (nn.GRU bi-directional 150->80, q_mu / q_logvar Linear 160->100, constructed in the
reference's order so that seeds give identical initial weights and state_dict keys).  The
arithmetic runs in the sm_100a kernels (csrc/gru.cu, gemm.cu) reached through RNN_VAE."""
import torch.nn as nn


def build_encoder(enc_type, **E_args):
    if enc_type != 'gru':
        raise ValueError('Please use GRU Encoder')
    return GRUEncoder(**E_args)


class GRUEncoder(nn.Module):
    def __init__(self, emb_dim, h_dim, z_dim, biGRU, layers, p_dropout):
        super().__init__()
        if not (emb_dim == 150 and h_dim == 80 and z_dim == 100 and biGRU and layers == 1):
            raise NotImplementedError(
                'cpg_b200 kernels are built for the reference geometry (emb 150, bi-GRU h 80, 1 layer, z 100); got '
                'emb_dim=%s h_dim=%s z_dim=%s biGRU=%s layers=%s' % (emb_dim, h_dim, z_dim, biGRU, layers))
        self.rnn = nn.GRU(input_size=emb_dim, hidden_size=h_dim, num_layers=layers, dropout=p_dropout,
                          bidirectional=biGRU, batch_first=True)
        self.biGRU_factor = 2 if biGRU else 1
        self.biGRU = biGRU
        self.q_mu = nn.Linear(self.biGRU_factor * h_dim, z_dim)
        self.q_logvar = nn.Linear(self.biGRU_factor * h_dim, z_dim)

    def forward(self, x):
        raise RuntimeError('GRUEncoder is evaluated through RNN_VAE.forward_encoder (token ids -> fused kernels); '
                           'feeding pre-embedded inputs is not part of the B200 hot path')
