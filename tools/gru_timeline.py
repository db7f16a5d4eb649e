"""Developer probe: clock64() timeline of one CTA / one step of the tcgen05 GRU forward kernels.
Needs a library built with -DCPG_GRU_TIMELINE (make NVFLAGS_EXTRA=-DCPG_GRU_TIMELINE); not part of the product."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'controlled-peptide-generation_b200'))
import torch
from cpg_b200 import engine, _lib
if os.environ.get('CPG_TL_LIB'): _lib._LIB_PATH = os.environ['CPG_TL_LIB']   # side build with -DCPG_GRU_TIMELINE
from oracle import wae as ow

dev = torch.device('cuda'); V, L, B = 24, 25, 4096
p = ow.random_params(V, seed=1)
st = engine.FlatState(V, dev); st.load(p)
tokens = ow.synthetic_tokens(B, V, seed=2).to(dev)
noise = engine.alloc_noise(B, L, dev)
engine.fill_step_noise(noise, 1, 0)
engine.train_step(st, tokens, noise, engine.make_hparams())
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 256)()
_lib.lib().cpg_debug_gru_timeline(buf)
names = ['pre-wait', 'mma-done-seen', 'phase1-done', 'after-bar', 'phase2-done', 'arrived', 'prefetch-issued']
for kid, which in enumerate(('fwd enc', 'fwd dec', 'bwd dec', 'bwd enc')):
    t = list(buf)[kid * 64:(kid + 1) * 64]
    base = min(x for x in t[:48] if x > 0)
    print(which, 'MMA start per chain', [t[0] - base, t[1] - base], 'setup', t[51] - t[50], 'total', t[52] - t[50])
    for ch in range(2):
        o = 8 + 16 * ch
        if t[o] > 0:
            print('  chain', ch, {n: t[o + i] - base for i, n in enumerate(names) if t[o + i] > 0})
