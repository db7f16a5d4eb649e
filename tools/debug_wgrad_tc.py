import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'controlled-peptide-generation_b200'))
import numpy as np, torch
from cpg_b200 import _lib
L_ = _lib.lib(); ctx = _lib.context()
dev = torch.device('cuda')
fn = L_.cpg_debug_wgrad_tc
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p] * 2 + [ctypes.c_int] + [ctypes.c_void_p] * 2 + [ctypes.c_int] * 2 + [ctypes.c_void_p] * 2 + [ctypes.c_int]

def run(HP, dg, hs, B, L, dbg=0):
    part = torch.full((148, 3 * HP, HP), -7.0, device=dev)
    ns = ctypes.c_int(0)
    rc = fn(ctx, None, HP, dg.data_ptr(), hs.data_ptr(), B, L, part.data_ptr(), ctypes.byref(ns), dbg)
    torch.cuda.synchronize()
    assert rc == 0, _lib.last_error()
    return part[:ns.value].sum(0).cpu(), ns.value

def expect(HP, dg, hs, B, L):
    n = B * L
    d = dg.view(n, 4, HP).cpu().double(); h = hs.view(n, HP).cpu().double()
    hp = torch.zeros_like(h); hp[1:] = h[:-1]
    valid = (torch.arange(n) % L != 0).double()[:, None]
    out = []
    for pl in (0, 1, 3):
        out.append((d[:, pl] * valid).t() @ hp)
    return torch.cat(out, 0)

for dbg in (0, 1, 2):
    B, L, HP = 2, 16, 80
    n = B * L
    dg, hs = torch.ones(n, 4 * HP), torch.ones(n, HP)
    got, ns = run(HP, dg.to(dev), hs.to(dev), B, L, dbg)
    print('dbg=%d ones: got[0,:4]=%s got[100,:4]=%s want 30 (31/32 with dbg&1)' % (dbg, got[0, :4].tolist(), got[100, :4].tolist()))
