"""Developer check on the GPU: ONE iteration from identical state under different library options (first configuration =
the reference one); prints the largest difference of every output / gradient / scalar and the per-kernel event times.
usage: python tools/lat_check.py B [B ...] -- "opt=val,opt=val" "opt=val" ...   (default: latent / RF tensor-core paths vs SIMT)"""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/controlled-peptide-generation_b200")
import torch
from cpg_b200 import engine, _lib
from oracle import wae as ow

dev = torch.device("cuda"); V = 24
_lib.set_option('gru_tensor_core', 2); _lib.set_option('dec_out_tensor_core', 2)
argv = sys.argv[1:]
confs = None
if '--' in argv:
    confs = [dict(kv.split('=') for kv in c.split(',') if kv) for c in argv[argv.index('--') + 1:]]
    argv = argv[:argv.index('--')]
confs = confs or [{'latent_tensor_core': 0, 'rf_tensor_core': 0}, {'latent_tensor_core': 2, 'rf_tensor_core': 0},
                  {'latent_tensor_core': 2, 'rf_tensor_core': 2}]
names = ['ref'] + ['c%d' % i for i in range(1, len(confs))]
for B in ([int(a) for a in argv] or [131, 4096]):
    p = ow.random_params(V, seed=11)
    tokens = ow.synthetic_tokens(B, V, seed=50).to(dev)
    noise = {k: v.to(dev).contiguous() for k, v in ow.draw_noise(B, seed=60).items()}
    res = {}
    for name, conf in zip(names, confs):
        for k, v in conf.items():
            _lib.set_option(k, int(v))
        st = engine.FlatState(V, dev); st.load(p)
        hp = engine.make_hparams(beta=1.0, lambda_logvar_l1=0.01, clip_norm=1e9)
        sc, ex = engine.train_step(st, tokens, noise, hp, want=('mu', 'logvar', 'z', 'logits'))
        torch.cuda.synchronize()
        out = {k: v.clone() for k, v in ex.items()}
        out.update({'g/' + k: v.clone() for k, v in st.views(st.grads).items()})
        out['scalars'] = sc.clone()
        res[name] = out
        _lib.profile_enable(True)
        for _ in range(3):
            st2 = engine.FlatState(V, dev); st2.load(p)
            engine.train_step(st2, tokens, noise, hp)
        rows_ = _lib.profile_read()
        _lib.profile_enable(False)
        t = {}
        for r in rows_:
            nm, ms, n = r[0], r[1], r[2]
            if 'latent' in nm or 'reparam' in nm or 'sgemm' in nm or 'prep' in nm or 'rf' in nm:
                t[nm] = '%.1f us x%d' % (1e3 * ms / max(n, 1), n // 3)
        print('B=%d %s %s:' % (B, name, conf), t)
    for name in names[1:]:
        print('B=%d %s vs ref' % (B, name))
        for k in res['ref']:
            a, b = res['ref'][k], res[name][k]
            d = (a - b).abs()
            i = int(d.argmax())
            print('  %-24s max|d| %.3e  max|a| %.3e  rel %.2e' % (k, float(d.max()), float(a.abs().max()), float(d.max() / (a.abs().max() + 1e-30))))
