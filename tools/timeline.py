"""Developer tool: timeline of ONE training iteration (eager launches, side stream on) from CUDA events around every
launch: start offset and duration per kernel and stream.  Shows which stream is the critical path.
usage: python tools/timeline.py [batch] [option=value ...]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'controlled-peptide-generation_b200'))
import torch
import bench
from cpg_b200 import engine, _lib, synth

dev = torch.device("cuda"); B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
for kv in sys.argv[2:]:
    k, v = kv.split('='); _lib.set_option(k, int(v))
cfg, model = bench.setup_model(dev)
st = model.bind_grads()
hp = engine.make_hparams(lr=cfg.vae.lr, z_regu=cfg.vae.z_regu_loss, mmd_sigma=cfg.losses.wae_mmd.sigma, rf_dim=cfg.losses.wae_mmd.rf_dim)
fs = engine.FusedStepper(st, B, 25, hp, seed=1, rf_dim=cfg.losses.wae_mmd.rf_dim)
tokens = synth.synthetic_tokens(B, bench.N_VOCAB, seed=2).to(dev)
for it in range(5):
    fs.step(tokens, it, 1.0)
torch.cuda.synchronize()
_lib.profile_enable(True)
for it in (5, 6):
    fs.step(tokens, it, 1.0)
    _lib.profile_read()
tl = _lib.profile_timeline()
_lib.profile_enable(False)
end = 0.0
for name, t0, dt in tl:
    end = max(end, t0 + dt)
    print('%8.1f us  +%6.1f us  %s%s' % (1e3 * t0, 1e3 * dt, '' if name.startswith('S0') else '                         ', name))
print('iteration span %.1f us, %d launches' % (1e3 * end, len(tl)))
