"""Developer probe of k_gru_bwd_fused: per-kernel time with parts of the work skipped (cpg_debug_bptt flags:
1 no dW_hh MMAs, 2 no dT MMAs, 4 no dh MMAs, 8 no gate prefetch, 16 no tile writes) and clock64 stamps of one step."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'controlled-peptide-generation_b200'))
import torch
from cpg_b200 import engine, _lib
if os.environ.get('CPG_TL_LIB'): _lib._LIB_PATH = os.environ['CPG_TL_LIB']   # side build with -DCPG_GRU_TIMELINE, synth

dev = torch.device('cuda')
V, L = 24, 25
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
st = engine.FlatState(V, dev)
torch.manual_seed(0)
st.params.copy_(torch.randn_like(st.params) * 0.1)
tokens = synth.synthetic_tokens(B, V, seed=2).to(dev)
noise = engine.alloc_noise(B, L, dev)
hp = engine.make_hparams()
L_ = _lib.lib()
L_.cpg_debug_bptt.argtypes = [ctypes.c_int, ctypes.c_void_p]
tl = torch.zeros(32, dtype=torch.int64, device=dev)
_lib.set_option('side_stream', 0)
def run(flags, with_tl=False):
    L_.cpg_debug_bptt(flags, ctypes.c_void_p(tl.data_ptr() if with_tl else 0))
    for i in range(2):
        engine.fill_step_noise(noise, 1238, i)
        engine.train_step(st, tokens, noise, hp)
    _lib.profile_enable(True)
    for i in range(5):
        engine.fill_step_noise(noise, 1238, i)
        engine.train_step(st, tokens, noise, hp)
    rows = {n: t / c for n, t, c in _lib.profile_read()}
    _lib.profile_enable(False)
    return rows
for flags in [int(x) for x in (sys.argv[2].split(',') if len(sys.argv) > 2 else '0,1,2,3,4,7,8,16,24,31'.split(','))]:
    r = run(flags)
    print('flags %2d: enc %.1f us  dec %.1f us' % (flags, 1e3 * r.get('k_gru_bwd_enc_fused', 0), 1e3 * r.get('k_gru_bwd_dec_fused', 0)))
for flags in (0, 8, 16, 24):
    run(flags, True)
    torch.cuda.synchronize()
    t = tl.cpu().tolist()
    for name, o in (('enc', 0), ('dec', 16)):
        m, e = t[o:o + 3], t[o + 8:o + 15]
        print('flags %2d %s | MMA warp: dh issue %d, dW issue %d | epilogue: wait dh %d, readout+bar %d, wait dW %d, math+stores %d, '
              'onehot+fence+arrive %d, prefetch issue %d | step (MMA start -> next arrive) %d'
              % (flags, name, m[1] - m[0], m[2] - m[1], e[1] - e[0], e[2] - e[1], e[3] - e[2], e[4] - e[3], e[5] - e[4],
                 e[6] - e[5], e[5] - m[0]))
L_.cpg_debug_bptt(0, ctypes.c_void_p(0))
