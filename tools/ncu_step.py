"""Short driver for ncu captures: a few fused WAE iterations at the bench configuration (B=4096),
then one CLaSS draw kernel and one beam-decode kernel.  Numbers printed under a profiler are not
bench values."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'controlled-peptide-generation_b200'))
import torch
from cpg_b200 import engine, sampling
from oracle import wae as ow
from oracle import cpu_baseline as cb

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = torch.device('cuda')
st = engine.FlatState(24, dev)
st.load(ow.random_params(24, seed=1))
tokens = ow.synthetic_tokens(B, 24, seed=2).to(dev)
noise = engine.alloc_noise(B, 25, dev)
hp = engine.make_hparams()
for i in range(steps):
    engine.fill_step_noise(noise, 1238, i)
    engine.train_step(st, tokens, noise, hp)
torch.cuda.synchronize()
if len(sys.argv) > 3:
    w, m, cv, clfs = cb.synthetic_class_setup()
    gmm = sampling.GmmDevice(w, m, cv, dev)
    spec = sampling.ClassifierSpec(clfs, dev)
    sampling.class_sample(gmm, spec, 10_000_000, 1)
    z = torch.randn(8192, 100, device=dev); c = torch.eye(2, device=dev)[torch.arange(8192, device=dev) % 2]
    sampling.beam_decode(st.params, 24, z, c)
    torch.cuda.synchronize()
print('done')
