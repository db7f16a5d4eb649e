"""Developer probe: encoder forward recurrence with and without the HBM stash stores."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'controlled-peptide-generation_b200'))
import torch
from cpg_b200 import engine, _lib
from oracle import wae as ow
dev = torch.device('cuda'); V, L, B = 24, 25, 4096
p = ow.random_params(V, seed=1)
st = engine.FlatState(V, dev); st.load(p)
tokens = ow.synthetic_tokens(B, V, seed=2).to(dev)
for i in range(3): engine.wae_encode(st.params, V, tokens)
torch.cuda.synchronize()
_lib.profile_enable(True)
for i in range(10): engine.wae_encode(st.params, V, tokens)
rows = _lib.profile_read(); _lib.profile_enable(False)
for name, t, n in rows: print('%-28s %8.3f ms x%d' % (name, t / max(n, 1), n))
