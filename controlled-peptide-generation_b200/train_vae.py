"""Phase-1 WAE/VAE training loop with the reference's entry point
`train_vae(cfgv, model, dataset)` (train_vae.py:13-68 of IBM/controlled-peptide-generation).

Default (cfg.b200.fused_step): each iteration is two calls into libcpg_b200 --
Philox noise, then forward -> the five losses -> BPTT -> clip_grad_norm_ -> Adam (with the
duplicated-embedding semantics of `vae_params()`), all on the current CUDA stream; the host only
moves the token batch (if the loader produced it on the host: one async copy from pinned memory into
a persistent device buffer) and, on logging iterations (or every `cfg.b200.sync_scalars_every`
iterations), reads the 16-float scalar block back (the reference reads nine `.item()`s every
iteration).  Under torch.distributed (one process per GPU) the batch each rank receives is its shard
of the global batch and the iteration runs through cpg_b200.parallel.dp_train_step.

cfg.b200.fused_step = False runs the reference-style loop (model() -> losses.* -> loss.backward()
-> clip_grad_norm_ -> torch.optim.Adam) over the same kernels through autograd.Functions.
"""
import sys

import torch
import torch.optim as optim

import cfg
import losses
import utils
from cpg_b200 import engine, parallel
from models.mutils import save_model
from tb_json_logger import log_value

train = None   # alias set below (BASELINE.json calls the entry point "train_vae.train()")
last_scalars = None   # host copy of the most recent scalar block read back (engine.SC slots)

_SCALAR_LOG = (('z_mu_L1', 'z_mu_l1'), ('z_logvar', 'z_logvar'), ('z_logvar_L1', 'logvar_l1'),
               ('z_logvar_KL_penalty', 'logvar_kl'), ('L_vae', 'loss'), ('L_vae_recon', 'recon'),
               ('L_vae_kl', 'kl'), ('L_wae_mmd', 'mmd'), ('L_wae_mmdrf', 'mmdrf'))


def _progress(rng):
    try:
        from tqdm import tqdm
        return tqdm(rng, disable=None), tqdm.write
    except Exception:  # noqa: BLE001
        return rng, print


def _log_sample(model, dataset, write):
    log_sent, _, _ = model.generate_sentences(1, sample_mode='categorical')
    write('Sample (cat T=1.0): "{}"'.format(dataset.idx2sentence(log_sent.squeeze())))
    sys.stdout.flush()


def _train_fused(cfgv, model, dataset):
    global last_scalars
    st = model.bind_grads()
    dev = st.device
    if not model.word_emb.weight.requires_grad:
        raise NotImplementedError('freeze_embeddings=True: the fused clip+Adam updates every VAE tensor; train with '
                                  'cfg.b200.fused_step = False (autograd path, optim.Adam over vae_params())')
    st.reset_optimizer()                             # the reference builds a fresh optim.Adam per call (train_vae.py:15)
    wm = cfg.losses.wae_mmd
    if wm.kernel != 'gaussian':
        raise NotImplementedError("only the gaussian MMD kernel is built")
    hp = engine.make_hparams(lr=cfgv.lr, clip_norm=cfgv.clip_grad, lambda_logvar_l1=cfgv.lambda_logvar_L1,
                             lambda_logvar_kl=cfgv.lambda_logvar_KL, z_regu=cfgv.z_regu_loss, mmd_sigma=wm.sigma,
                             rf_dim=wm.rf_dim, compute_full_mmd=True)
    p_word, p_out = model.decoder.word_dropout.p, model.decoder.p_out_dropout
    seed = int(cfg.b200.noise_seed)
    every = max(1, int(cfg.b200.full_mmd_every))
    sync_every = int(getattr(cfg.b200, 'sync_scalars_every', 0))
    distributed = parallel.is_distributed()
    if distributed:
        parallel.sync_replicas(st)                   # rank 0's weights everywhere; moments / step were just reset
        # rank-distinct noise rows; rf_w / rf_b come from the shared seed inside alloc_noise
        rank_seed = (seed + 0x9E3779B97F4A7C15 * (1 + parallel.dist.get_rank())) % (1 << 63)
    stepper, global_batch = None, None
    # data parallel: (compute_full_mmd, B, L, settings) -> parallel.GraphedDPStepper, kept on the state across calls (the
    # captured graph is valid as long as the flat buffers it reads are: they belong to this state)
    dp_steppers = st.__dict__.setdefault('_dp_steppers', {})
    it_range, write = _progress(range(cfgv.s_iter, cfgv.s_iter + cfgv.n_iter + 1))
    last_it = cfgv.s_iter + cfgv.n_iter
    # Host batches go to the device one iteration ahead, on a copy stream, into one of two buffers: the H2D copy of
    # batch i+1 runs under the kernels of iteration i (same number of next_batch calls as the reference loop).
    copy_stream = torch.cuda.Stream(device=dev)
    read_stream = torch.cuda.Stream(device=dev)     # device -> host reads of the scalar block, off the compute stream
    bufs, used_ev = [None, None], [None, None]       # device token buffers / "last reader has been enqueued" events
    staged = None                                     # (host tokens, slot or None, copy event or None) of the next iteration
    pinned_scal, pending_read = None, None

    def stage(tokens, slot):
        if engine._lib._on_device(tokens):
            return tokens, None, None
        Bn, Ln = tokens.shape
        if bufs[slot] is None or bufs[slot].shape != (Bn, Ln):
            bufs[slot] = torch.empty(Bn, Ln, dtype=torch.int64, device=dev)
            used_ev[slot] = None
            # the caching allocator may hand back a block whose last readers are still queued on the compute
            # stream (a previous token buffer, noise of a replaced stepper): order the copy behind them
            copy_stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(copy_stream):
            if used_ev[slot] is not None:
                copy_stream.wait_event(used_ev[slot])                    # its previous reader (two iterations ago)
            bufs[slot].copy_(tokens, non_blocking=True)                  # async when the loader pins its batches
            ev = copy_stream.record_event()
        return tokens, slot, ev

    for n_done, it in enumerate(it_range):
        log_it = it % cfgv.cheaplog_every == 0 or it % cfgv.expsvlog_every == 0
        if staged is None:
            staged = stage(dataset.next_batch('train_vae').text, n_done % 2)
        tokens, slot, ev = staged
        staged = None
        B, L = tokens.shape
        if stepper is None or stepper.B != B or stepper.L != L:
            stepper = engine.FusedStepper(st, B, L, hp, seed=seed, p_word=p_word, p_out=p_out, rf_dim=wm.rf_dim)
            global_batch = parallel.global_batch_size(B, dev) if distributed else B
            if distributed:
                parallel.assert_distinct_shards(tokens.to(dev))          # warns when every rank feeds the same batch
        if slot is None:
            tok = tokens if tokens.is_contiguous() else tokens.contiguous()
        else:
            torch.cuda.current_stream(dev).wait_event(ev)
            tok = bufs[slot]
        beta = float(utils.anneal(cfgv.beta, it))
        hp.compute_full_mmd = 1 if (it % every == 0 or log_it) else 0
        if distributed:
            # one captured graph per rank (noise + both phases + the collectives + clip/Adam); the log-only full-kernel MMD is
            # evaluated on some iterations only, so there is one stepper (and graph) per value of that switch
            hp_k = type(hp).from_buffer_copy(hp)
            hp_k.adam_step, hp_k.beta, hp_k.global_batch = 0, 0.0, 0          # per-step / per-call fields are not part of the key
            mode = str(getattr(cfg.b200, 'dp_full_mmd', 'local'))
            key = (B, L, rank_seed, p_word, p_out, mode, int(global_batch), bytes(hp_k))
            ds = dp_steppers.get(key)
            if ds is None:
                ds = parallel.GraphedDPStepper(st, B, L, type(hp).from_buffer_copy(hp), stepper.noise, rank_seed, global_batch,
                                               p_word=p_word, p_out=p_out, full_mmd=mode,
                                               graph=bool(getattr(cfg.b200, 'dp_graph', True)))
                dp_steppers[key] = ds
            scal = ds.step(tok, it, beta)
        else:
            scal = stepper.step(tok, it, beta)
        if slot is not None:
            used_ev[slot] = torch.cuda.current_stream(dev).record_event()
        if it != last_it:                                                # next batch: fetch + H2D under this iteration
            staged = stage(dataset.next_batch('train_vae').text, (n_done + 1) % 2)
        # device -> host read of the scalar block.  Iterations that print need it now; the periodic read
        # (cfg.b200.sync_scalars_every) is pipelined: copied to pinned memory asynchronously and collected after the
        # NEXT iteration has been enqueued, so the GPU never idles behind the host round trip.
        if pending_read is not None:
            pending_read[1].synchronize()
            last_scalars = pending_read[0].clone()
            pending_read = None
        if log_it:
            last_scalars = vals = scal.cpu()
        elif sync_every > 0 and it % sync_every == 0:
            if pinned_scal is None:
                pinned_scal = torch.empty(scal.shape, dtype=scal.dtype, pin_memory=True)
            # on its own stream, behind this iteration: the next iteration's launch does not queue behind the copy (the fused
            # stepper alternates between two scalar blocks, so the block being read is not rewritten for a whole iteration)
            # (the data-parallel graph owns ONE scalar block: there the copy stays on the compute stream)
            if distributed:
                pinned_scal.copy_(scal, non_blocking=True)
                pending_read = (pinned_scal, torch.cuda.current_stream(dev).record_event())
            else:
                read_stream.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(read_stream):
                    pinned_scal.copy_(scal, non_blocking=True)
                    pending_read = (pinned_scal, read_stream.record_event())
        if log_it:
            for name, slot in _SCALAR_LOG:
                log_value('train_' + name, float(vals[engine.SC[slot]]), it)
            log_value('train_beta', beta, it)
            write('ITER {} TRAINING (phase 1). loss_vae: {:.4f}; loss_recon: {:.4f}; loss_kl: {:.4f}; '
                  'loss_mmd: {:.4f}; Grad_norm: {:.4e} '.format(
                      it, float(vals[engine.SC['loss']]), float(vals[engine.SC['recon']]),
                      float(vals[engine.SC['kl']]), float(vals[engine.SC['mmd']]),
                      float(vals[engine.SC['grad_norm']])))
            _log_sample(model, dataset, write)
        if it % cfgv.expsvlog_every == 0 and it > 0:
            save_model(model, cfgv.chkpt_path.format(it))
    if pending_read is not None:
        pending_read[1].synchronize()
        last_scalars = pending_read[0].clone()
    return st


def _train_modular(cfgv, model, dataset):
    trainer = optim.Adam(model.vae_params(), lr=cfgv.lr)
    it_range, write = _progress(range(cfgv.s_iter, cfgv.s_iter + cfgv.n_iter + 1))
    for it in it_range:
        log_it = it % cfgv.cheaplog_every == 0 or it % cfgv.expsvlog_every == 0
        inputs = dataset.next_batch('train_vae')
        tokens = inputs.text
        if not engine._lib._on_device(tokens):
            tokens = tokens.to(model._param_device())
        beta = utils.anneal(cfgv.beta, it)
        (z_mu, z_logvar), (z, c), dec_logits = model(tokens, q_c='prior', sample_z=1)
        recon_loss = losses.recon_dec(tokens, dec_logits)
        kl_loss = losses.kl_gaussianprior(z_mu, z_logvar)
        # the full-kernel MMD is differentiated only when it is the regulariser (same draw order as the reference)
        wae_mmd_loss = losses.wae_mmd_gaussianprior(z if cfgv.z_regu_loss == 'mmd' else z.detach(), method='full_kernel')
        wae_mmdrf_loss = losses.wae_mmd_gaussianprior(z, method='rf')
        z_regu_loss = {'kl': kl_loss, 'mmd': wae_mmd_loss, 'mmdrf': wae_mmdrf_loss}[cfgv.z_regu_loss]
        z_logvar_L1 = z_logvar.abs().sum(1).mean(0)
        z_logvar_KL_penalty = losses.kl_gaussian_sharedmu(z_mu, z_logvar)
        loss = recon_loss + beta * z_regu_loss + cfgv.lambda_logvar_L1 * z_logvar_L1 \
            + cfgv.lambda_logvar_KL * z_logvar_KL_penalty
        trainer.zero_grad()
        loss.backward()
        grad_norm = torch.nn.utils.clip_grad_norm_(model.vae_params(), cfgv.clip_grad)
        trainer.step()
        if log_it:
            for name, val in (('z_mu_L1', z_mu.data.abs().mean()), ('z_logvar', z_logvar.data.mean()),
                              ('z_logvar_L1', z_logvar_L1), ('z_logvar_KL_penalty', z_logvar_KL_penalty),
                              ('L_vae', loss), ('L_vae_recon', recon_loss), ('L_vae_kl', kl_loss),
                              ('L_wae_mmd', wae_mmd_loss), ('L_wae_mmdrf', wae_mmdrf_loss)):
                log_value('train_' + name, val.item(), it)
            log_value('train_beta', beta, it)
            write('ITER {} TRAINING (phase 1). loss_vae: {:.4f}; loss_recon: {:.4f}; loss_kl: {:.4f}; '
                  'loss_mmd: {:.4f}; Grad_norm: {:.4e} '.format(it, loss.item(), recon_loss.item(), kl_loss.item(),
                                                               wae_mmd_loss.item(), float(grad_norm)))
            _log_sample(model, dataset, write)
        if it % cfgv.expsvlog_every == 0 and it > 0:
            save_model(model, cfgv.chkpt_path.format(it))


def train_vae(cfgv, model, dataset):
    print('Training base vae ...')
    if bool(cfg.b200.fused_step):
        return _train_fused(cfgv, model, dataset)
    return _train_modular(cfgv, model, dataset)


train = train_vae
