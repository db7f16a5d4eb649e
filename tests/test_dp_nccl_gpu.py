"""Two real GPUs, NCCL: the data-parallel iteration (phase1 -> all-reduce -> phase2 -> all-reduce ->
clip+Adam) on two shards equals the single-GPU fused iteration on the whole batch, and the replicas
stay bit-identical.  Skipped on boxes with fewer than 2 GPUs."""
import os
import socket

import numpy as np
import pytest
import torch

from helpers import assert_params_close
from oracle import wae as ow

pytestmark = pytest.mark.gpu
V, B = 24, 70


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import sys
    from conftest import PKG
    sys.path.insert(0, PKG)
    import torch.distributed as dist
    from cpg_b200 import engine, parallel
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    p = ow.random_params(V, seed=5)
    tokens = ow.synthetic_tokens(B, V, seed=6)
    noise = ow.draw_noise(B, seed=7)
    lo, hi = parallel.shard_bounds(B, rank, world)
    st = engine.FlatState(V, dev)
    st.load(p)
    nz = {k: (v[lo:hi] if v.shape[0] == B else v).to(dev).contiguous() for k, v in noise.items()}
    out = []
    for it in range(2):
        hp = engine.make_hparams(beta=1.0 + 0.5 * it)
        sc = parallel.dp_train_step(st, tokens[lo:hi].to(dev).contiguous(), nz, hp, full_mmd='global')
        out.append(sc.cpu())
    flat = st.params.clone()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        ret['scalars'] = torch.stack(out)
        ret['params'] = flat.cpu()
        ret['replicas_equal'] = all(torch.equal(gathered[0], g) for g in gathered[1:])
    dist.destroy_process_group()


def test_two_gpu_nccl_equals_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    from cpg_b200 import engine
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret['replicas_equal']
    dev = torch.device('cuda', 0)
    p = ow.random_params(V, seed=5)
    tokens = ow.synthetic_tokens(B, V, seed=6).to(dev)
    noise = {k: v.to(dev) for k, v in ow.draw_noise(B, seed=7).items()}
    st = engine.FlatState(V, dev)
    st.load(p)
    for it in range(2):
        sc, _ = engine.train_step(st, tokens, noise, engine.make_hparams(beta=1.0 + 0.5 * it))
        got = ret['scalars'][it]
        for k in ('loss', 'recon', 'kl', 'mmd', 'mmdrf', 'logvar_kl', 'grad_norm'):
            assert float(got[engine.SC[k]]) == pytest.approx(float(sc[engine.SC[k]]), rel=1e-4, abs=1e-7), (it, k)
    single = st.views(st.params)
    multi = st.views(ret['params'].to(dev))
    assert_params_close(multi, single, st.views(st.grads), 'dp2', max_outliers=3)
