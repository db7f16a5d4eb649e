"""World-size-2 gloo test (CPU) of the PRODUCT's data-parallel iteration, cpg_b200.parallel.dp_train_step:
start-up broadcast, the asynchronous all-reduce of the coupled statistics, the single gradient all-reduce
that also carries the NLL sum, replicated clip+Adam, the ragged all-gather of the global MMD, and the
replica checks -- with the CUDA phases replaced by their oracle restatement (oracle/dp.py: OracleEngine),
since kernels cannot run here.  Result == the single-process oracle on the whole batch.  The same function
with the real kernels is covered on the GPU box by tests/test_dp_nccl_gpu.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dp as odp
from oracle import wae as ow

V, B = 24, 13          # ragged split: 7 + 6


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    import sys
    import warnings
    from conftest import PKG
    sys.path.insert(0, PKG)
    from cpg_b200 import engine, parallel
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    p = ow.random_params(V, seed=4 + 17 * rank)               # replicas start DIFFERENT ...
    st = odp.OracleState(p)
    st.params = torch.cat([st.p[k].reshape(-1) for k in ow.UNIQUE_VAE_PARAMS])
    st.adam_m, st.adam_v = torch.zeros_like(st.params), torch.zeros_like(st.params)
    diverged = False
    if world > 1:
        try:
            parallel.check_replicas(st)
        except RuntimeError:
            diverged = True
    parallel.sync_replicas(st)                                # ... and rank 0's weights win
    parallel.check_replicas(st)
    off = 0
    for k in ow.UNIQUE_VAE_PARAMS:
        n = st.p[k].numel()
        st.p[k] = st.params[off:off + n].view_as(st.p[k]).clone()
        off += n
    st.p['decoder.emb.weight'] = st.p['word_emb.weight']
    tokens = ow.synthetic_tokens(B, V, seed=8)
    noise = ow.draw_noise(B, seed=9)
    lo, hi = parallel.shard_bounds(B, rank, world)
    sl = {k: (v[lo:hi] if v.shape[0] == B else v) for k, v in noise.items()}
    with warnings.catch_warnings(record=True) as wrn:
        warnings.simplefilter('always')
        same_ok = parallel.assert_distinct_shards(tokens)                 # every rank holds the SAME batch: must warn
        distinct_ok = parallel.assert_distinct_shards(tokens[lo:hi])
    hp = engine.make_hparams(beta=1.3)
    sc = parallel.dp_train_step(st, tokens[lo:hi], sl, hp, full_mmd='global', eng=odp.OracleEngine())
    flat = torch.cat([st.p[k].reshape(-1) for k in ow.UNIQUE_VAE_PARAMS])
    both = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(both, flat)
    if rank == 0:
        ret.update(scalars=sc, grads=st.grads.clone(), params=flat, replicas_equal=all(torch.equal(both[0], b) for b in both),
                   diverged_before_sync=diverged, warned=(not same_ok) and len(wrn) >= 1, distinct_ok=distinct_ok,
                   global_batch=hp.global_batch, step=st.step)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_dp_train_step_equals_full_batch():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret['diverged_before_sync'] and ret['replicas_equal'] and ret['warned'] and ret['distinct_ok']
    assert ret['global_batch'] == B and ret['step'] == 1
    p = ow.random_params(V, seed=4)
    tokens = ow.synthetic_tokens(B, V, seed=8)
    noise = ow.draw_noise(B, seed=9)
    scal, grads, _ = ow.train_step(p, {}, tokens, noise, beta=1.3)
    S = odp.OracleEngine.SC
    sc = ret['scalars']
    for k in ('loss', 'recon', 'kl', 'mmd', 'mmdrf', 'logvar_kl', 'grad_norm'):
        assert float(sc[S[k]]) == pytest.approx(scal[k], rel=1e-4, abs=1e-7), k
    want = torch.cat([grads[k].reshape(-1) for k in ow.UNIQUE_VAE_PARAMS])
    assert float((ret['grads'] - want).abs().max()) <= 2e-5 * float(want.abs().max()) + 1e-8
    wantp = torch.cat([p[k].reshape(-1) for k in ow.UNIQUE_VAE_PARAMS])
    assert float((ret['params'] - wantp).abs().max()) < 2e-5


def test_shard_bounds_and_ragged_gather_single_process():
    import sys
    from conftest import PKG
    sys.path.insert(0, PKG)
    from cpg_b200.parallel import shard_bounds
    for n, w in ((13, 2), (4096, 8), (5, 8), (100, 3)):
        b = [shard_bounds(n, r, w) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 1
