"""World-size-2 gloo test (CPU) of the data-parallel exchange logic: shard the batch, all-reduce the
coupled statistics, all-reduce the gradients, and compare with the single-process oracle on the whole
batch.  The per-rank math is the oracle restatement of the two CUDA phases (oracle/dp.py); the
collectives, shard bounds and reduction order are the product's (cpg_b200.parallel)."""
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
    from conftest import PKG
    sys.path.insert(0, PKG)
    from cpg_b200.parallel import shard_bounds
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    p = ow.random_params(V, seed=4)
    tokens = ow.synthetic_tokens(B, V, seed=8)
    noise = ow.draw_noise(B, seed=9)
    lo, hi = shard_bounds(B, rank, world)
    sl = {k: (v[lo:hi] if v.shape[0] == B else v) for k, v in noise.items()}
    coupled = odp.phase1_coupled(p, tokens[lo:hi], sl)
    dist.all_reduce(coupled)                                   # exchange 1
    gsize = torch.tensor([hi - lo])
    dist.all_reduce(gsize)
    grads, nll = odp.phase2_local_grads(p, tokens[lo:hi], sl, coupled, int(gsize), beta=1.3)
    flat = torch.cat([grads[k].reshape(-1) for k in ow.UNIQUE_VAE_PARAMS])
    dist.all_reduce(flat)                                      # exchange 2
    nll_t = torch.tensor([nll])
    dist.all_reduce(nll_t)
    if rank == 0:
        ret['flat'] = flat
        ret['recon'] = float(nll_t / coupled[0])
        ret['coupled'] = coupled
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_shards_equal_full_batch():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    p = ow.random_params(V, seed=4)
    tokens = ow.synthetic_tokens(B, V, seed=8)
    noise = ow.draw_noise(B, seed=9)
    scal, grads, _ = ow.train_step({k: v.clone() for k, v in p.items()}, {}, tokens, noise, beta=1.3,
                                   with_full_mmd=False)
    want = torch.cat([grads[k].reshape(-1) for k in ow.UNIQUE_VAE_PARAMS])
    got = ret['flat']
    assert float((got - want).abs().max()) <= 2e-5 * float(want.abs().max()) + 1e-8
    assert ret['recon'] == pytest.approx(scal['recon'], rel=1e-5)
    c = ret['coupled']
    assert float(c[2]) / B == pytest.approx(scal['kl'], rel=1e-5)
    assert float(c[3]) / B == pytest.approx(scal['logvar_kl'], rel=1e-5)
    d = (c[8:508] - c[508:]) / B
    assert float((d ** 2).sum()) == pytest.approx(scal['mmdrf'], rel=1e-4)
