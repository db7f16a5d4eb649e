"""Developer probe: where the warp roles of k_wgrad_tc (CTA 0) spend their cycles.  Needs -DCPG_GRU_TIMELINE."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'controlled-peptide-generation_b200'))
import torch
from cpg_b200 import engine, _lib
from oracle import wae as ow
dev = torch.device('cuda'); V, L, B = 24, 25, 4096
p = ow.random_params(V, seed=1)
st = engine.FlatState(V, dev); st.load(p)
tokens = ow.synthetic_tokens(B, V, seed=2).to(dev)
noise = engine.alloc_noise(B, L, dev)
_lib.set_option('side_stream', 0)
for i in range(3):
    engine.fill_step_noise(noise, 1, i)
    engine.train_step(st, tokens, noise, engine.make_hparams())
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 32)()
_lib.lib().cpg_debug_wgrad_timeline(buf)
t = list(buf)
print('last k_wgrad_tc launch (decoder), CTA 0, cycles:')
print('  producer : wait empty %d, issue TMA %d, total %d, stages %d' % (t[0], t[1], t[2], t[3]))
print('  MMA      : wait converted %d, issue+commit %d, total %d' % (t[4], t[5], t[6]))
print('  converter: wait full %d, convert %d, total %d ; epilogue %d' % (t[8], t[9], t[10], t[11]))
