"""DEVELOPER TOOL: point the cpg_b200 binding at the g++-built CPU emulation of the
kernels (tools/cuda_emu/libcpg_emu.so, `make -C .../csrc emu`) so that kernel
indexing and host plumbing can be debugged in the GPU-less build container.

Never imported by the product, by tests/, by bench.py or by __graft_entry__.py:
parity claims are made on the GPU only.  Usage (from a scratch script):

    import tools.cuda_emu.use_emu  # before touching cpg_b200
"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, 'controlled-peptide-generation_b200'))
from cpg_b200 import _lib  # noqa: E402

_lib._LIB_PATH = os.path.join(ROOT, 'tools', 'cuda_emu', 'libcpg_emu.so')
_lib._require_cuda = lambda: None
_lib._device_index = lambda device=None: 0
_lib.tensor_device = lambda device=None: torch.device('cpu')
_lib.stream_ptr = lambda: ctypes.c_void_p(None)
_lib._on_device = lambda t: True
