"""Developer check of latent_tc.cu on the GPU: ONE iteration from identical state with the fp32 SIMT dense layers
(latent_tensor_core=0) and with the tcgen05 ones (64- and 128-row tiles); prints the largest difference of every
output / gradient and the per-kernel event times of the kernels involved."""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/controlled-peptide-generation_b200")
import torch
from cpg_b200 import engine, _lib
from oracle import wae as ow

dev = torch.device("cuda"); V = 24
_lib.set_option('gru_tensor_core', 2); _lib.set_option('dec_out_tensor_core', 2)
for B in ([int(a) for a in sys.argv[1:]] or [131, 4096]):
    p = ow.random_params(V, seed=11)
    tokens = ow.synthetic_tokens(B, V, seed=50).to(dev)
    noise = {k: v.to(dev).contiguous() for k, v in ow.draw_noise(B, seed=60).items()}
    res = {}
    for name, flag, rows in (('simt', 0, 64), ('tc64', 2, 64), ('tc128', 2, 128)):
        _lib.set_option('latent_tensor_core', flag); _lib.set_option('latent_tile_rows', rows)
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
            if 'latent' in nm or 'reparam' in nm or 'sgemm' in nm or 'prep_weights' in nm:
                t[nm] = '%.1f us x%d' % (1e3 * ms / max(n, 1), n // 3)
        print('B=%d %s:' % (B, name), t)
    for name in ('tc64', 'tc128'):
        print('B=%d %s vs simt' % (B, name))
        for k in res['simt']:
            a, b = res['simt'][k], res[name][k]
            d = (a - b).abs()
            i = int(d.argmax())
            print('  %-24s max|d| %.3e  max|a| %.3e  rel %.2e' % (k, float(d.max()), float(a.abs().max()), float(d.max() / (a.abs().max() + 1e-30))))
