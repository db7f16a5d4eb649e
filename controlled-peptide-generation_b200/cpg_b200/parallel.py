"""Data-parallel WAE iteration: one process per GPU, batch sharded by sample, and the three exchange
points of SURVEY.md 8(e) done with torch.distributed (NCCL over NVLink on the GPU box, gloo in the
CPU tests of the host logic):

  phase 1 (local)  forward through the decoder GRU + local statistics
  all-reduce SUM   `coupled` = [n_tok, -, 5 latent sums, RF feature sums of z, of z_prior]  (~4 KB)
  phase 2 (local)  CE with the GLOBAL token count, RF-MMD gradient from the GLOBAL feature means,
                   BPTT, weight gradients (each rank's share of the global-batch gradient)
  all-reduce SUM   flat gradient buffer (1.03 MB)
  clip + Adam      identical on every rank (weights stay replicated bit-for-bit)

The reference has no distributed code; the contract is "N ranks on shards == one process on the
concatenated batch".  The log-only full-kernel MMD is evaluated on the local shard unless
`full_mmd='global'` (all-gather of z and z_prior, then every rank computes the global value).
"""
import torch
import torch.distributed as dist

from . import engine


def is_distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shard_bounds(n, rank, world):
    """Contiguous, balanced split of n items: ranks < n % world get one extra."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_batch_size(local_batch, device, group=None):
    sizes = torch.tensor([local_batch], device=device, dtype=torch.int64)
    dist.all_reduce(sizes, group=group)
    return int(sizes.item())


def dp_train_step(state, tokens, noise, hp, p_out=0.3, group=None, full_mmd='local', global_batch=None):
    """One data-parallel iteration on this rank's shard.  Returns the scalar block (device); after the
    call scalars hold GLOBAL recon / KL / RF-MMD values (the kernels compose them from the reduced
    statistics), the full-kernel MMD slot is local unless full_mmd == 'global'.  Pass `global_batch`
    (sum of the shard sizes) when it is constant to avoid one tiny all-reduce + host sync per step."""
    world = dist.get_world_size(group)
    hp.global_batch = int(global_batch) if global_batch else global_batch_size(tokens.shape[0], tokens.device, group)
    state.step += 1
    hp.adam_step = state.step
    local_noise = noise
    z_prior_full = noise.get('z_prior_full')
    if full_mmd != 'local':
        local_noise = dict(noise)
        local_noise.pop('z_prior_full', None)            # phase 2 must not compute the local value
    coupled, z = engine.step_phase1(state, tokens, local_noise, hp, p_out)
    dist.all_reduce(coupled, group=group)
    scalars = engine.step_phase2(state, tokens, local_noise, hp, coupled, p_out)
    # recon needs the global sum of NLL: the local sum sits in SC_NLL_SUM
    dist.all_reduce(state.grads, group=group)
    gn = engine.clip_adam(state, hp)
    fix = scalars[engine.SC['nll_sum']:engine.SC['nll_sum'] + 1].clone()
    dist.all_reduce(fix, group=group)
    ntok = scalars[engine.SC['ntok']]
    recon_global = fix[0] / torch.clamp(ntok, min=1.0)
    delta = recon_global - scalars[engine.SC['recon']]
    scalars[engine.SC['recon']] = recon_global
    scalars[engine.SC['loss']] += delta
    scalars[engine.SC['grad_norm']] = gn[0]
    if full_mmd == 'global' and z_prior_full is not None:
        zs = [torch.empty_like(z) for _ in range(world)]
        zp = [torch.empty_like(z_prior_full) for _ in range(world)]
        dist.all_gather(zs, z.contiguous(), group=group)
        dist.all_gather(zp, z_prior_full.contiguous(), group=group)
        scalars[engine.SC['mmd']] = engine.mmd_full(torch.cat(zs), torch.cat(zp), hp.mmd_sigma)[0]
    return scalars
