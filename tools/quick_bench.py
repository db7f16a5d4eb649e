"""Developer tool: ms/step of the fused iteration (captured graph) at B=4096 for sets of library options.
usage: python tools/quick_bench.py ["opt=val,opt=val" ...]   (each argument = one configuration; '' = defaults)"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'controlled-peptide-generation_b200'))
import torch
import bench
from cpg_b200 import engine, _lib, synth
if os.environ.get('CPG_TL_LIB'): _lib._LIB_PATH = os.environ['CPG_TL_LIB']   # side build (A/B runs)

dev = torch.device("cuda"); B = 4096
cfg, model = bench.setup_model(dev)
st = model.bind_grads()
hp = engine.make_hparams(lr=cfg.vae.lr, z_regu=cfg.vae.z_regu_loss, mmd_sigma=cfg.losses.wae_mmd.sigma, rf_dim=cfg.losses.wae_mmd.rf_dim)
tokens = synth.synthetic_tokens(B, bench.N_VOCAB, seed=2).to(dev)
DEFAULTS = {}
for conf in (sys.argv[1:] or ['']):
    opts = dict(kv.split('=') for kv in conf.split(',') if kv)
    for k, v in opts.items():
        _lib.set_option(k, int(v))
    fs = engine.FusedStepper(st, B, 25, hp, seed=1, rf_dim=cfg.losses.wae_mmd.rf_dim)
    for it in range(10):
        fs.step(tokens, it, 1.0)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for it in range(100):
            fs.step(tokens, 10 + it, 1.0)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 100)
    print('%-50s %.4f ms/step' % (conf or '(defaults)', best), flush=True)
