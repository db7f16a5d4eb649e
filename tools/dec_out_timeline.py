"""Developer probe: cycles per phase of k_dec_out_tc (CTA 0, thread 0; each phase includes its closing barrier).
Needs a library built with -DCPG_GRU_TIMELINE."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'controlled-peptide-generation_b200'))
import torch
from cpg_b200 import engine, _lib
if os.environ.get('CPG_TL_LIB'): _lib._LIB_PATH = os.environ['CPG_TL_LIB']   # side build with -DCPG_GRU_TIMELINE
from oracle import wae as ow
dev = torch.device('cuda'); V, L, B = 24, 25, 4096
st = engine.FlatState(V, dev); st.load(ow.random_params(V, seed=1))
tokens = ow.synthetic_tokens(B, V, seed=2).to(dev)
noise = engine.alloc_noise(B, L, dev)
_lib.set_option('side_stream', 0)
for i in range(3):
    engine.fill_step_noise(noise, 1, i)
    engine.train_step(st, tokens, noise, engine.make_hparams())
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 16)()
_lib.lib().cpg_debug_dec_out_timeline(buf)
t = list(buf)
print('k_dec_out_tc CTA 0, cycles summed over its tiles:')
for name, i in (('tile start (prev tile tail)', 7), ('S1 stage hd', 0), ('S2 transpose', 1), ('M1 + E1 softmax', 2), ('M2 + E2 dh', 3)):
    print('  %-28s %8d' % (name, t[i]))

import numpy as np
cb = (ctypes.c_ulonglong * (3 * 160))()
if hasattr(_lib.lib(), 'cpg_debug_dec_out_ctas'):
    _lib.lib().cpg_debug_dec_out_ctas(cb)
    t = np.array(list(cb), dtype=np.float64).reshape(160, 3)
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    print('CTAs %d: start spread %.1f us; set-up mean %.1f us; run (start -> exit) mean %.1f us max %.1f us; last exit at %.1f us after the first start'
          % (len(t), (t[:, 0].max() - t0) / 1e3, (t[:, 1] - t[:, 0]).mean() / 1e3, (t[:, 2] - t[:, 0]).mean() / 1e3,
             (t[:, 2] - t[:, 0]).max() / 1e3, (t[:, 2].max() - t0) / 1e3))
