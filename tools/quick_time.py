"""Developer timing probe: per-kernel CUDA-event breakdown of the fused WAE step and the CLaSS kernels."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'controlled-peptide-generation_b200'))
import torch, numpy as np
from cpg_b200 import engine, sampling, _lib
from oracle import wae as ow

dev = torch.device('cuda')
V, L = 24, 25
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
p = ow.random_params(V, seed=1)
st = engine.FlatState(V, dev); st.load(p)
tokens = ow.synthetic_tokens(B, V, seed=2).to(dev)
noise = engine.alloc_noise(B, L, dev)
hp = engine.make_hparams()
SIDE = int(sys.argv[2]) if len(sys.argv) > 2 else 1      # 0: everything on one stream (per-kernel times are then each kernel's own)
def step(i):
    engine.fill_step_noise(noise, 1238, i, overlap=True)
    return engine.train_step(st, tokens, noise, hp)
for i in range(3): step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 20
e0.record()
for i in range(K): s, _ = step(3 + i)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print('B=%d step %.3f ms -> %.0f seq/s ; loss %.4f' % (B, ms, B / ms * 1e3, float(s[0])))
_lib.set_option('side_stream', SIDE)
_lib.profile_enable(True)
for i in range(5): step(30 + i)
rows = _lib.profile_read()
_lib.profile_enable(False)
_lib.set_option('side_stream', 1)
tot = sum(r[1] for r in rows)
for name, t, n in sorted(rows, key=lambda r: -r[1]):
    print('  %-28s %8.3f ms/step  x%-3d %5.1f%%' % (name, t / 5, n // 5, 100 * t / tot))
print('  sum of kernels %.3f ms/step' % (tot / 5))
# CLaSS
rs = np.random.RandomState(0); K_ = 100
w = rs.dirichlet(np.ones(K_)); m = rs.randn(K_, 100) * 0.5; cv = rs.uniform(0.05, 0.2, (K_, 100))
gmm = sampling.GmmDevice(w, m, cv, dev)
clfs = [('amp', rs.randn(100).astype(np.float32) * 0.1, np.float32(0.1), 1), ('tox', rs.randn(100).astype(np.float32) * 0.1, np.float32(-0.2), 0)]
spec = sampling.ClassifierSpec(clfs, dev)
n = 10_000_000
for want_z in (True, False):
    out = sampling.class_sample(gmm, spec, n, 1, want_z=want_z, want_scores=want_z)
    torch.cuda.synchronize()
    e0.record(); out = sampling.class_sample(gmm, spec, n, 2, want_z=want_z, want_scores=want_z); e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1)
    print('class_sample n=%d want_z=%s: %.3f ms -> %.3g draws/s, acc rate %.4f' % (n, want_z, t, n / t * 1e3, float(out['n_accepted']) / n))
nz = 8192
z = torch.randn(nz, 100, device=dev); c = torch.eye(2, device=dev)[torch.arange(nz, device=dev) % 2]
sampling.beam_decode(st.params, V, z, c); torch.cuda.synchronize()
e0.record(); sampling.beam_decode(st.params, V, z, c); e1.record(); torch.cuda.synchronize()
print('beam decode n=%d: %.3f ms -> %.3g seq/s' % (nz, e0.elapsed_time(e1), nz / e0.elapsed_time(e1) * 1e3))
e0.record(); sampling.gmm_logpdf(gmm, z); e1.record(); torch.cuda.synchronize()
print('gmm_logpdf n=%d: %.3f ms' % (nz, e0.elapsed_time(e1)))
