"""Developer probe: clock64() timeline of one CTA / one step of the tcgen05 GRU forward kernels.
Needs a library built with -DCPG_GRU_TIMELINE (make NVFLAGS_EXTRA=-DCPG_GRU_TIMELINE); not part of the product."""
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
engine.fill_step_noise(noise, 1, 0)
for which in ('dec', 'enc'):
    # the decoder forward runs after the encoder forward, so the stamps left are the decoder's; for the
    # encoder use the encoder-only entry
    if which == 'dec':
        engine.train_step(st, tokens, noise, engine.make_hparams())
    else:
        engine.wae_encode(st.params, V, tokens)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 64)()
    _lib.lib().cpg_debug_gru_timeline(buf)
    t = list(buf)
    base = min(x for x in t if x > 0)
    names = ['pre-wait', 'mma-done-seen', 'phase1-done', 'after-bar', 'phase2-done', 'arrived']
    print(which, 'MMA: start', [t[0] - base, t[1] - base], 'issued', [t[2] - base, t[3] - base])
    print('  kernel: setup', t[51] - t[50], 'total', t[52] - t[50], '; phase-2 stamps (item: math+X done, stores done):',
          [(t[40 + 2 * i] - base, t[41 + 2 * i] - base) for i in range(2)])
    for sub in range(2):
        for w in range(2):
            o = 8 + 16 * sub + 8 * w
            if t[o] > 0:
                print('  sub', sub, 'warp', 'first' if w == 0 else 'last', {n: t[o + i] - base for i, n in enumerate(names)})
