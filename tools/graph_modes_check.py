"""Developer check: the captured-graph iteration equals eager launches bit for bit for every latent regulariser and with the
logged full-kernel MMD on / off (B = 4096, Philox noise)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'controlled-peptide-generation_b200'))
import torch
from cpg_b200 import engine, _lib, synth
from oracle import wae as ow

dev = torch.device('cuda'); V, B = 24, int(sys.argv[1]) if len(sys.argv) > 1 else 4096
tokens = synth.synthetic_tokens(B, V, seed=2).to(dev)
p = ow.random_params(V, seed=3)
ok = True
for z_regu in ('mmdrf', 'kl', 'mmd'):
    for full in (1, 0):
        if z_regu == 'mmd' and not full:
            continue
        res = []
        for graph in (1, 0):
            _lib.set_option('cuda_graph', graph)
            st = engine.FlatState(V, dev); st.load(p)
            hp = engine.make_hparams(z_regu=z_regu, lambda_logvar_l1=0.01)
            hp.compute_full_mmd = full
            fs = engine.FusedStepper(st, B, 25, hp, seed=11)
            sc = [fs.step(tokens, it, 0.3 + 0.1 * it).clone() for it in range(5)]
            torch.cuda.synchronize()
            res.append((torch.stack(sc).cpu(), st.params.clone().cpu()))
        same = torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
        fin = bool(torch.isfinite(res[0][0]).all())
        ok &= same and fin
        print('%-6s full_mmd=%d  graph == eager: %s  finite: %s  loss %.5f' % (z_regu, full, same, fin, float(res[0][0][-1, 0])))
_lib.set_option('cuda_graph', 1)
print('ALL OK' if ok else 'MISMATCH')
