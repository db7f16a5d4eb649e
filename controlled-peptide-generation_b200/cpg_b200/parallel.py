"""Data-parallel WAE iteration: one process per GPU, batch sharded by sample, exchanges done with
torch.distributed (NCCL over NVLink on the GPU box; gloo in the CPU tests of this host logic).

  phase 1 (local)     forward through the decoder GRU; the local statistics that couple the batch
                      (`coupled` = [n_tok, -, 5 latent sums, RF feature sums of z, of z_prior], ~4 KB) are
                      produced on the library's side stream while the decoder recurrence runs
  all-reduce #1 SUM   of `coupled`, ASYNC, in two pieces on the library's internal streams: the token count right
                      after the token preparation, the latent / RF sums under the decoder recurrence; the main
                      stream never waits for them directly (phase 2 orders the consumers)
  phase 2 (local)     CE with the GLOBAL token count, RF-MMD gradient from the GLOBAL feature means, BPTT,
                      weight gradients (this rank's share of the global-batch gradient)
  all-reduce #2 SUM   flat gradient (1.03 MB) + an 8-float tail carrying the local NLL sum (and the
                      partial sums of the distributed full-kernel MMD) -- one collective, no third one
  clip + Adam         identical on every rank (weights stay replicated bit-for-bit)

The reference has no distributed code; the contract is "N ranks on shards == one process on the
concatenated batch" (SURVEY.md 8e).  Replicas are made identical at start-up (`sync_replicas`: broadcast
of params / Adam moments / step from rank 0) and can be verified at any time (`check_replicas`).
Each rank's loader must yield ITS shard of the global batch (`assert_distinct_shards` warns when two
ranks present the same token batch).

The log-only full-kernel MMD: full_mmd='local' evaluates it on the local shard (cheap; a per-shard
value), 'global' all-gathers z / z_prior (ragged shards are padded) and evaluates the global-batch value.
"""
import warnings

import torch
import torch.distributed as dist

from . import engine as _engine


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


def sync_replicas(state, group=None, src=0):
    """Start-up broadcast: parameters, Adam moments and the step count of rank `src` to every rank."""
    for t in (state.params, state.adam_m, state.adam_v):
        dist.broadcast(t, src, group=group)
    step = torch.tensor([state.step], device=state.params.device, dtype=torch.int64)
    dist.broadcast(step, src, group=group)
    state.step = int(step.item())
    return state


def check_replicas(state, group=None):
    """Raise if the replicas' parameters / Adam moments / step differ (bit-level checksums, min == max)."""
    def chk(t):
        b = t.detach().contiguous().view(torch.int32).to(torch.int64)
        return torch.stack([b.sum(), (b * (torch.arange(b.numel(), device=b.device) % 8191 + 1)).sum()])
    sums = torch.cat([chk(state.params), chk(state.adam_m), chk(state.adam_v),
                      torch.tensor([state.step], device=state.params.device, dtype=torch.int64)])
    lo, hi = sums.clone(), sums.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    if not torch.equal(lo, hi):
        raise RuntimeError('data-parallel replicas have diverged (params / Adam state / step checksums differ '
                           'across ranks); call sync_replicas() after loading weights')


def assert_distinct_shards(tokens, group=None):
    """Same-seed loaders would train N copies of one batch: warn when two ranks hold identical tokens."""
    t = tokens.detach().to(torch.int64).reshape(-1)
    w = torch.arange(t.numel(), device=t.device, dtype=torch.int64) % 1021 + 1
    h = torch.stack([(t * w).sum(), t.sum()]).reshape(1, 2)
    world = dist.get_world_size(group)
    hs = [torch.empty_like(h) for _ in range(world)]
    dist.all_gather(hs, h, group=group)
    keys = [tuple(x.reshape(-1).tolist()) for x in hs]
    if len(set(keys)) < world:
        warnings.warn('data-parallel ranks received identical token batches: give every rank its own shard of the '
                      'global batch (rank-sharded loader / distinct sampler seeds)')
        return False
    return True


def all_gather_rows(t, group=None, even=False):
    """Concatenation over ranks of [n_r, ...] tensors.  even=True: every rank holds the same n (one collective, no
    host synchronisation); else the sizes are exchanged first and the rows padded (ragged shards)."""
    world = dist.get_world_size(group)
    if even:
        out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous(), group=group)
        return out
    return all_gather_ragged(t, group)


def all_gather_ragged(t, group=None):
    """Concatenation over ranks of [n_r, ...] tensors whose n_r may differ (padded exchange)."""
    world = dist.get_world_size(group)
    n = torch.tensor([t.shape[0]], device=t.device, dtype=torch.int64)
    ns = [torch.empty_like(n) for _ in range(world)]
    dist.all_gather(ns, n, group=group)
    ns = [int(x.item()) for x in ns]
    nmax = max(ns)
    pad = t.contiguous()
    if pad.shape[0] < nmax:
        pad = torch.cat([pad, pad.new_zeros((nmax - pad.shape[0],) + tuple(pad.shape[1:]))])
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[:k] for o, k in zip(outs, ns)])


def dp_train_step(state, tokens, noise, hp, p_out=0.3, group=None, full_mmd='local', global_batch=None, eng=None,
                  _advance=True):
    """One data-parallel iteration on this rank's shard.  Returns the scalar block (device): recon / KL /
    RF-MMD / loss are GLOBAL values; the full-kernel MMD slot is the local-shard value unless
    full_mmd == 'global'.  Pass `global_batch` (sum of the shard sizes) when it is constant to avoid one
    tiny all-reduce + host sync per step.  `eng` (default: cpg_b200.engine) provides the local phases."""
    eng = eng or _engine
    hp.global_batch = int(global_batch) if global_batch else global_batch_size(tokens.shape[0], tokens.device, group)
    if _advance:
        state.step += 1
        hp.adam_step = state.step
    local_noise = noise
    z_prior_full = noise.get('z_prior_full')
    if full_mmd != 'local':
        local_noise = dict(noise)
        local_noise.pop('z_prior_full', None)            # phase 2 must not compute the local value
    gext = state.grads_ext
    tail = gext[state.grads.numel():]
    coupled, z = eng.step_phase1(state, tokens, local_noise, hp, p_out)
    # exchange 1, asynchronous and off the main stream: the token count on the lane that produced it right after the token
    # preparation (phase 2's decoder-output layer waits for that lane), the latent / RF sums behind the lane that produced
    # them under the decoder recurrence (the RF-MMD chain of phase 2 continues on that lane)
    with eng.ntok_stream(tokens.device):
        dist.all_reduce(coupled[0:1], group=group, async_op=True).wait()     # stream-level wait (no host block with NCCL)
    with eng.stats_stream(tokens.device):
        dist.all_reduce(coupled[1:], group=group, async_op=True).wait()
    scalars = eng.step_phase2(state, tokens, local_noise, hp, coupled, p_out)
    eng.dp_pack_tail(tail)
    dist.all_reduce(gext, group=group)                   # exchange 2: gradients + NLL sum
    g = eng.SC['grad_norm']
    eng.clip_adam(state, hp, out=scalars[g:g + 1])
    eng.dp_apply_tail(tail, scalars)
    if full_mmd == 'global' and z_prior_full is not None:
        world = dist.get_world_size(group)
        even = hp.global_batch == world * tokens.shape[0]
        zs = all_gather_rows(z, group, even)
        zp = all_gather_rows(z_prior_full, group, even)
        m = eng.SC['mmd']
        eng.mmd_full(zs, zp, hp.mmd_sigma, out=scalars[m:m + 1])
    return scalars


_LIVE_STEPPERS = None


def release_graphs():
    """Drop every captured data-parallel graph of this process (see GraphedDPStepper.release)."""
    for ds in list(_LIVE_STEPPERS or ()):
        ds.release()


def _track(stepper):
    """Remember the stepper and make dist.destroy_process_group() release the captured graphs first: a live graph that
    holds captured collectives keeps the communicator busy at teardown (the process would hang at exit)."""
    global _LIVE_STEPPERS
    if _LIVE_STEPPERS is None:
        import weakref
        _LIVE_STEPPERS = weakref.WeakSet()
        inner = dist.destroy_process_group

        def destroy_process_group(*a, **kw):
            release_graphs()
            return inner(*a, **kw)
        dist.destroy_process_group = destroy_process_group
    _LIVE_STEPPERS.add(stepper)


class GraphedDPStepper:
    """The perf-mode data-parallel iteration (Philox noise + dp_train_step, collectives included) as ONE captured CUDA graph
    per rank, replayed every step: the ~35 kernels, the three all-reduces and the fork / join of the library's lanes cost one
    graph launch on the host instead of ~45 eager launches from Python.  The per-step scalars (beta, Adam bias corrections,
    noise counter) live in the library's device block, refreshed by one tiny kernel before each replay
    (engine.step_dyn_write).  Two eager iterations come first (allocations, attribute set-up); if the capture fails
    (e.g. a process group whose collectives cannot be captured) the stepper stays on eager launches -- same results.

    `tokens` is copied into a static buffer each step (the graph reads fixed addresses)."""

    def __init__(self, state, B, L, hp, noise, seed, global_batch, group=None, p_word=0.3, p_out=0.3, full_mmd='local',
                 graph=True):
        self.state, self.hp, self.noise, self.seed = state, hp, noise, int(seed)
        self.group, self.p_word, self.p_out, self.full_mmd, self.global_batch = group, p_word, p_out, full_mmd, int(global_batch)
        self.B, self.L = B, L
        dev = state.params.device
        self.dev = dev
        self.tokens = torch.zeros(B, L, dtype=torch.int64, device=dev)
        self.want_graph = bool(graph) and full_mmd == 'local'
        self.graph, self.scalars, self.calls = None, None, 0
        self.force_eager = False                          # e.g. while the per-kernel profiler is on (it sees eager launches only)
        _track(self)

    def _body(self, it):
        _engine.fill_step_noise(self.noise, self.seed, it, self.p_word, self.p_out, overlap=True)
        return dp_train_step(self.state, self.tokens, self.noise, self.hp, p_out=self.p_out, group=self.group,
                             full_mmd=self.full_mmd, global_batch=self.global_batch, _advance=False)

    def step(self, tokens, it, beta):
        if tokens.data_ptr() != self.tokens.data_ptr():
            self.tokens.copy_(tokens, non_blocking=True)
        st, hp = self.state, self.hp
        st.step += 1
        hp.adam_step = st.step
        hp.beta = float(beta)
        self.calls += 1
        if self.graph is None and self.want_graph and self.calls >= 3 and not self.force_eager:
            self._capture(it)
        if self.graph is not None and not self.force_eager:
            _engine.step_dyn_write(self.dev, hp, it)
            self.graph.replay()
            return self.scalars
        return self._body(it)

    def release(self):
        """Drop the captured graph (and the scalar block it owns).  Call before dist.destroy_process_group(): a live graph
        that holds captured collectives keeps the communicator busy at teardown."""
        if self.graph is not None:
            torch.cuda.synchronize(self.dev)
            self.graph, self.scalars = None, None
            import gc
            gc.collect()
            torch.cuda.synchronize(self.dev)
            self.calls = 0                                # (a later step() captures again after two eager iterations)

    def _capture(self, it):
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        try:
            _engine.step_dyn_write(self.dev, self.hp, it)
            _engine.step_dyn_use(self.dev, True)
            with torch.cuda.graph(g, capture_error_mode='thread_local'):
                sc = self._body(it)
            self.graph, self.scalars = g, sc
        except Exception as e:                                   # noqa: BLE001 -- any capture failure: stay eager
            warnings.warn('data-parallel step: CUDA-graph capture failed (%s: %s); using eager launches'
                          % (type(e).__name__, e))
            self.want_graph = False
        finally:
            _engine.step_dyn_use(self.dev, False)
        torch.cuda.synchronize(self.dev)
