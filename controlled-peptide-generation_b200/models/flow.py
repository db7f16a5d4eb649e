"""Normalising flows on the latent code with the reference's classes and parameter names (models/flow.py of
IBM/controlled-peptide-generation: build_flow :7-15, PlanarFlow :30-60, RadialFlow :63-96, AlternatingFlow :99-160).
`forward(z, train=True)` -> (z', loss) in train mode, z' otherwise, like the reference; the transform and its
log-determinant loss run in ONE kernel launch over all layers (cpg_flow_forward).  Inference / evaluation only: the
kernel is not differentiated (flow > 0 is outside the reference's training default, cfg.py:267).
"""
import math
from ctypes import c_float, c_int, c_void_p

import numpy as np
import torch
import torch.nn as nn

from cpg_b200 import _lib

PLANAR, RADIAL = 0, 1


def build_flow(flow_type, flow_layers, z_dim):
    if flow_type == 'planar':
        return PlanarFlow(flow_layers, z_dim)
    if flow_type == 'radial':
        return RadialFlow(flow_layers, z_dim)
    if flow_type == 'alternating':
        return AlternatingFlow(flow_layers, z_dim)
    raise ValueError('Please use either planar, radial, or alternating flow.')


class Flow(nn.Module):
    def __init__(self, flow_layers, z_dim):
        super().__init__()
        if z_dim != 100 or not 1 <= flow_layers <= 16:
            raise NotImplementedError('the flow kernel is built for z_dim = 100 and 1..16 layers')
        self.flow, self.z_dim = flow_layers, z_dim

    def _planar_params(self):
        self.planar_weight = nn.ParameterList([nn.Parameter(torch.empty(1, self.z_dim).uniform_(-0.01, 0.01)) for _ in range(self.flow)])
        self.planar_bias = nn.ParameterList([nn.Parameter(torch.empty(1).uniform_(-0.01, 0.01)) for _ in range(self.flow)])
        self.planar_scale = nn.ParameterList([nn.Parameter(torch.empty(1, self.z_dim).uniform_(-0.01, 0.01)) for _ in range(self.flow)])

    def _radial_params(self):
        self.radial_initial = nn.ParameterList([nn.Parameter(torch.empty(1, self.z_dim).uniform_(-0.01, 0.01)) for _ in range(self.flow)])
        self.radial_alpha = nn.ParameterList([nn.Parameter(torch.empty(1).uniform_(0.01, 1.0)) for _ in range(self.flow)])
        self.radial_beta = nn.ParameterList([nn.Parameter(torch.empty(1).uniform_(-0.01, 0.01)) for _ in range(self.flow)])

    def layer_kinds(self):
        raise NotImplementedError

    # "Maintain invertibility" (reference flow.py:44-48 / :77-79): scalar checks that may move a parameter in place
    def _maintain_planar(self, i):
        w, s = self.planar_weight[i], self.planar_scale[i]
        margin = float((s.data * w.data).sum())
        if margin < -1:
            component = -1 + math.log(1 + math.e ** margin) - margin
            s.data += component * w.data / w.data.norm(2)

    def _maintain_radial(self, i):
        a, b = self.radial_alpha[i], self.radial_beta[i]
        if float(b.data) < -float(a.data):
            b.data = -a.data + torch.log(1 + math.e ** b.data)

    def forward(self, z, train=True):
        kinds = self.layer_kinds()
        n = len(kinds)
        va, vb, sa, sb, keep = [], [], [], [], []
        dev = z.device
        for i, k in enumerate(kinds):
            if k == PLANAR:
                self._maintain_planar(i)
                w = self.planar_weight[i].data.reshape(-1).contiguous().float()
                s = self.planar_scale[i].data.reshape(-1).contiguous().float()
                keep += [w, s]
                va.append(w.data_ptr()); vb.append(s.data_ptr())
                sa.append(float(self.planar_bias[i].data)); sb.append(float((w * s).sum()))
            else:
                self._maintain_radial(i)
                z0 = self.radial_initial[i].data.reshape(-1).contiguous().float()
                keep.append(z0)
                va.append(z0.data_ptr()); vb.append(0)
                sa.append(float(self.radial_alpha[i].data)); sb.append(float(self.radial_beta[i].data))
        zin = z.detach().contiguous().float()
        zout = torch.empty_like(zin)
        loss = torch.zeros(1, device=dev)
        L = _lib.lib()
        _lib.check(L.cpg_flow_forward(_lib.context(dev), _lib.stream_ptr(), _lib.ptr(zin), zin.shape[0], n,
                                      (c_int * n)(*kinds), (c_void_p * n)(*va), (c_void_p * n)(*vb),
                                      (c_float * n)(*sa), (c_float * n)(*sb), 1 if train else 0, _lib.ptr(zout),
                                      _lib.ptr(loss)), 'cpg_flow_forward')
        zout.flowed = True
        return (zout, loss[0]) if train else zout


class PlanarFlow(Flow):
    def __init__(self, flow_layers, z_dim):
        super().__init__(flow_layers, z_dim)
        self._planar_params()

    def layer_kinds(self):
        return [PLANAR] * self.flow


class RadialFlow(Flow):
    def __init__(self, flow_layers, z_dim):
        super().__init__(flow_layers, z_dim)
        self._radial_params()

    def layer_kinds(self):
        return [RADIAL] * self.flow


class AlternatingFlow(Flow):
    """Planar on even layers, radial on odd ones; both parameter sets exist for every layer (reference :99-125)."""
    def __init__(self, flow_layers, z_dim):
        super().__init__(flow_layers, z_dim)
        self._planar_params()
        self._radial_params()

    def layer_kinds(self):
        return [PLANAR if i % 2 == 0 else RADIAL for i in range(self.flow)]
