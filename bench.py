#!/usr/bin/env python
"""Benchmark of the hot path named by BASELINE.json: WAE phase-1 training throughput (seq/s) on
synthetic peptide batches, config "WAE phase-1 full config, batch=4096 len<=25, 1xB200, fp32"
(per-GPU batch 4096 under weak scaling for N > 1), plus the CLaSS accepted-samples/s figure.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One JSON line on stdout (rank 0).  `value` = device-resident fused iteration (inputs in HBM);
`e2e` = the same metric through the reference-facing call train_vae.train_vae(cfgv, model, dataset)
with the token batch in pinned HOST memory (H2D every step) and the scalar block read back every
step.  `roofline` = the dominant kernel, timed live with CUDA events on the launching stream.
`cpu_baseline` = the reference's CPU path on this box's host cores (bounded sample).
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'controlled-peptide-generation_b200')
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

N_VOCAB, SEQ_LEN, BATCH = 24, 25, 4096
METRIC, UNIT = 'wae_train_seq_per_s', 'seq/s'
WORKLOAD = 'WAE phase-1 full config, batch=4096 len<=25, fp32 (BASELINE.json configs[1])'


def load_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu summary."""
    path = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    try:
        with open(path) as f:
            return json.load(f).get(kernel)
    except (OSError, ValueError):
        return None


def load_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fh:
            pk = json.load(fh)
        return pk, 'measured'
    except Exception:  # noqa: BLE001
        return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(',')]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:  # noqa: BLE001
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for j, n in enumerate(names) if any(r[3 + j].lower().startswith('active') for r in self.rows)]
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(self.rows[0][1]), 'reasons': reasons,
                'samples': len(self.rows), 'power_w_max': max(float(r[2]) for r in self.rows)}


# per-kernel algorithmic work of ONE launch at batch B: (flops, bytes)
def kernel_work(B, L=SEQ_LEN, V=N_VOCAB, R=500):
    He, Hd, HP, Z = 80, 102, 104, 100
    f4 = 4
    fwd_enc = (2 * 2 * B * L * He * 3 * He, 2 * B * L * (5 * He) * f4)           # writes h + (r, z, n, hn), both directions
    fwd_dec = (2 * B * L * Hd * 3 * Hd, B * L * (5 * HP) * f4)
    bwd_enc = (2 * 2 * B * L * He * 3 * He, 2 * B * L * (5 * He + 4 * He) * f4)  # reads 4 gate planes + h_prev, writes 4 dg planes
    bwd_dec = (2 * B * L * Hd * 3 * Hd, B * L * (6 * HP + 4 * HP) * f4)          # + dh_out
    # the twelve small dense products of one iteration (heads, [z;c] projection, RF map and their transposes): (M, N, K)
    gemms = [(B, Z, 2 * He)] * 2 + [(B, 3 * HP, HP)] + [(B, R, Z)] * 2 + [(B, HP, 3 * HP), (3 * HP, HP, B)] + \
            [(B, 2 * He, Z)] * 2 + [(Z, 2 * He, B)] * 2 + [(B, Z, R)]
    g_flops = sum(2 * m * n * k for m, n, k in gemms)
    g_bytes = sum((m * k + k * n + m * n) * f4 for m, n, k in gemms)
    return {
        'k_gru_fwd_enc': fwd_enc, 'k_gru_fwd_enc_tc': fwd_enc,
        'k_gru_fwd_dec': fwd_dec, 'k_gru_fwd_dec_tc': fwd_dec,
        'k_gru_bwd_enc': bwd_enc, 'k_gru_bwd_enc_tc': bwd_enc,
        'k_gru_bwd_dec': bwd_dec, 'k_gru_bwd_dec_tc': bwd_dec,
        'k_wgrad_hh_enc': (2 * B * L * He * 3 * He, B * L * (4 * He) * f4),
        'k_wgrad_hh_dec': (2 * B * L * Hd * 3 * Hd, B * L * (4 * HP) * f4),
        # one pass over the 4 dg planes + h: dW_hh (3 planes x H) and the token-table gradient (4 planes x 32 one-hot columns)
        'k_wgrad_tc_enc': (2 * B * L * (3 * He * He + 4 * He * 32), B * L * (5 * He) * f4),
        'k_wgrad_tc_dec': (2 * B * L * (3 * HP * HP + 4 * HP * 32), B * L * (5 * HP) * f4),
        'k_mmd_gram_tc': (3 * 2 * B * B * 100, 2 * B * 128 * f4),
        'k_mmd_gram_tc2': (3 * 2 * B * B * 100, 2 * B * 128 * f4),
        'k_dec_out': (2 * 3 * B * L * Hd * V, B * L * (2 * HP * f4 + Hd)),
        'k_dec_out_tc': (2 * 3 * B * L * Hd * V, B * L * (2 * HP * f4 + Hd)),
        'k_mmd_gram': (3 * 2 * B * B * 100, 2 * B * 100 * f4),
        'k_dtable': (B * L * 4 * HP, B * L * 4 * HP * f4),
        'k_sgemm': (g_flops / len(gemms), g_bytes / len(gemms)),                  # average of the 12 launches of one iteration
        'k_input_grads': (2 * V * 150 * (2 * 3 * He + 3 * HP), V * (2 * 4 * He + 4 * HP) * f4),
    }


def setup_model(device):
    import numpy as np
    import torch
    import cfg
    from models.model import RNN_VAE
    torch.manual_seed(cfg.seed)
    np.random.seed(cfg.seed)
    model = RNN_VAE(n_vocab=N_VOCAB, max_seq_len=cfg.max_seq_len, **cfg.model).to(device)
    return cfg, model


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    from cpg_b200 import _lib, engine, parallel, sampling
    from oracle import wae as ow
    from oracle import cpu_baseline as cb

    cfg, model = setup_model(dev)
    st = model.bind_grads()
    B, L, K, W = args.batch, SEQ_LEN, args.steps, args.warmup
    tokens_host = ow.synthetic_tokens(B, N_VOCAB, seed=1238 + rank).pin_memory()
    tokens = tokens_host.to(dev)
    noise = engine.alloc_noise(B, L, dev, seed=cfg.seed)
    hp = engine.make_hparams(lr=cfg.vae.lr, z_regu=cfg.vae.z_regu_loss, mmd_sigma=cfg.losses.wae_mmd.sigma,
                             rf_dim=cfg.losses.wae_mmd.rf_dim)
    seed = cfg.b200.noise_seed + 7919 * rank
    gb = B * world
    counter = {'it': 0}

    def step():
        it = counter['it']
        counter['it'] += 1
        hp.beta = float(ow.anneal_beta(it))
        engine.fill_step_noise(noise, seed, it, overlap=True)      # next reader is the train step below
        if world > 1:
            return parallel.dp_train_step(st, tokens, noise, hp, global_batch=gb)
        return engine.train_step(st, tokens, noise, hp)[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(W, 3)):
        step()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    if clocks:
        clocks.start()
    l0 = _lib.launch_count()
    total_ms = timed(step, K)
    launches = _lib.launch_count() - l0
    ms_per_step = total_ms / K
    value = gb * K / (total_ms / 1e3)

    # ---- per-kernel durations (CUDA events around every launch of the same step, same stream)
    # The step overlaps its latency-bound loss / weight-gradient kernels with the recurrences on a side stream;
    # for the per-kernel pass everything is serialised on one stream so that each duration is the kernel's own.
    _lib.set_option('side_stream', 0)
    _lib.profile_enable(True)
    prof_ms = timed(step, K)
    rows = _lib.profile_read()
    _lib.profile_enable(False)
    _lib.set_option('side_stream', 1)
    per_kernel = {name: (ms / max(cnt, 1), cnt / K) for name, ms, cnt in rows}
    step_share = {name: ms / K for name, ms, cnt in rows}
    # dominant kernel = largest share among single-launch kernels with a work model (k_sgemm is 12 different small
    # products, listed in per_kernel with its average)
    work_names = set(kernel_work(B).keys()) - {'k_sgemm'}
    cand = {k: v for k, v in step_share.items() if k in work_names} or step_share
    dom = max(cand, key=cand.get)
    peaks, peak_src = load_peaks()
    work = kernel_work(B)
    flops, nbytes = work.get(dom, (0, 0))
    dur_s = per_kernel[dom][0] / 1e3
    tf = flops / dur_s / 1e12 if dur_s > 0 else 0.0
    gbs = nbytes / dur_s / 1e9 if dur_s > 0 else 0.0
    tensor_frac = tf / peaks['bf16_tflops_sustained']
    hbm_frac = gbs / peaks['hbm_gbs']
    if tensor_frac >= hbm_frac:
        roof = {'bound': 'tensor', 'achieved': round(tf, 3), 'peak': peaks['bf16_tflops_sustained'], 'unit': 'TFLOP/s',
                'frac': round(tensor_frac, 5)}
    else:
        roof = {'bound': 'hbm', 'achieved': round(gbs, 1), 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                'frac': round(hbm_frac, 5)}
    per_kernel_roof = {}
    for name, (kms, per_step) in per_kernel.items():
        if name in work and kms > 0:
            f, nb = work[name]
            per_kernel_roof[name] = {'ms': round(kms, 4), 'launches_per_step': per_step,
                                     'tflops': round(f / (kms / 1e3) / 1e12, 2), 'gbs': round(nb / (kms / 1e3) / 1e9, 1),
                                     'frac_tensor': round(f / (kms / 1e3) / 1e12 / peaks['bf16_tflops_sustained'], 4),
                                     'frac_hbm': round(nb / (kms / 1e3) / 1e9 / peaks['hbm_gbs'], 4)}
    roof.update({'per_kernel': per_kernel_roof, 'kernel': dom, 'kernel_ms': round(per_kernel[dom][0], 4), 'launches_per_step': per_kernel[dom][1],
                 'share_of_step': round(step_share[dom] / sum(step_share.values()), 4), 'traffic': load_traffic(dom),
                 'peak_source': peak_src + ' (MEASURED_PEAKS.json, sustained)' if peak_src == 'measured' else peak_src,
                 'note': 'dominant kernel = largest share of the step by CUDA events; per_kernel lists every kernel with a work model '
                         '(FLOP/s against the bf16 tensor peak, bytes against the HBM copy peak); traffic = dram bytes of one launch from '
                         'the committed ncu --set full capture (profiles/), null if that kernel was not captured',
                 'top_kernels_ms_per_step': {k: round(v, 4) for k, v in sorted(step_share.items(), key=lambda x: -x[1])[:8]}})

    # ---- end to end through the reference-facing API: host tokens -> train_vae.train_vae -> host scalars
    import tb_json_logger
    import train_vae as tv
    cfgv = cfg.Bunch(cfg.vae)
    cfgv.update(cfg.shared)
    cfgv.cheaplog_every, cfgv.expsvlog_every = 10 ** 9, 10 ** 9    # no sample generation / checkpoints in the timed loop
    cfg.b200.sync_scalars_every = 1                                # ... but the scalar block is read back every iteration
    ds = types.SimpleNamespace(next_batch=lambda name: types.SimpleNamespace(text=tokens_host),
                               idx2sentence=lambda idxs, print_special_tokens=True: '')
    tb_json_logger.configure()

    def e2e_run(k):
        cfgv.s_iter, cfgv.n_iter = 10 ** 6, k - 1                  # beta at its final value; k iterations
        with contextlib.redirect_stdout(io.StringIO()):
            tv.train_vae(cfgv, model, ds)
    e2e_run(max(W, 3))
    barrier()
    t0 = time.perf_counter()
    e2e_run(K)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = gb * K / float(e2e_s.item())
    clock_summary = clocks.summary() if clocks else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- CLaSS (secondary metric of BASELINE.json): accepted samples/s, Philox draws + scores + accept
    w, means, covs, clfs = cb.synthetic_class_setup()
    gmm = sampling.GmmDevice(w, means, covs, dev)
    spec = sampling.ClassifierSpec(clfs, dev)
    n_draws = 10_000_000
    sampling.class_sample(gmm, spec, n_draws, 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = sampling.class_sample(gmm, spec, n_draws, 2)
    e1.record()
    torch.cuda.synchronize()
    cms = e0.elapsed_time(e1)
    acc = int(out['n_accepted'].item())
    zb = torch.randn(8192, 100, device=dev)
    cbm = torch.eye(2, device=dev)[torch.arange(8192, device=dev) % 2]
    sampling.beam_decode(st.params, N_VOCAB, zb, cbm)
    torch.cuda.synchronize()
    e0.record()
    sampling.beam_decode(st.params, N_VOCAB, zb, cbm)
    e1.record()
    torch.cuda.synchronize()
    class_block = {'metric': 'class_accepted_samples_per_s', 'value': acc / (cms / 1e3), 'unit': 'samples/s',
                   'draws_per_s': n_draws / (cms / 1e3), 'accept_rate': acc / n_draws, 'n_draws': n_draws,
                   'bytes_written_per_draw': 425, 'hbm_gbs': 425 * n_draws / (cms / 1e3) / 1e9,
                   'beam_decode_seq_per_s': 8192 / (e0.elapsed_time(e1) / 1e3)}

    # ---- CPU baseline on this box's host cores (bounded sample)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cb.time_wae_cpu(B, N_VOCAB, steps=5, warmup=1, budget_s=90.0)      # ~13 s of CPU work on 16 host threads
        cpu = {'value': r['seq_per_s'], 'unit': UNIT, 'cores': r['cores'], 'kind': r['kind'],
               'sample': '%d full iterations at batch %d after 1 warm-up (%.0f ms/step)' % (
                   r['steps_timed'], B, r['ms_per_step'])}
        rc = cb.time_class_cpu(1_000_000)
        class_block['cpu_baseline'] = {'value': rc['accepted_per_s'], 'unit': 'samples/s', 'cores': 1, 'kind': rc['kind'],
                                       'sample': 'rejection_sample(1e6): %.2f s, accept rate %.3f' % (rc['seconds'], rc['accept_rate'])}
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': max(W, 3),
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'fp32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'per_gpu_batch': B, 'global_batch': gb, 'seq_len': L, 'n_vocab': N_VOCAB,
                   'parallelism': 'dp%d' % world if world > 1 else 'single',
                   'l2_policy': 'per-step working set (activation stash ~1.1 GB) exceeds the 126 MB L2; no flush needed',
                   'noise': 'Philox in-kernel, regenerated every step',
                   'arithmetic': 'fp32 storage and accumulation; recurrence / decoder-output contractions as split-bf16 '
                                 '(x1+x2, 3 products; logits 3 terms) tcgen05 MMAs, weight-gradient and MMD Gram contractions tf32, '
                                 'dense layers fp32 SIMT'},
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': B * L * 8, 'd2h_bytes_per_step': 16 * 4 + 8,
                'api': 'train_vae.train_vae(cfgv, model, dataset): pinned host tokens copied H2D every step (one step ahead, copy stream), '
                  'scalar block copied D2H every step (collected after the next step is enqueued)'},
        'gpu_launches': launches, 'launches_per_step': launches / K,
        'profiled_ms_per_step': prof_ms / K,
        'roofline': roof, 'cpu_baseline': cpu, 'clocks': clock_summary, 'class': class_block,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    """The reference's own CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from oracle import cpu_baseline as cb
    K, W = args.steps, max(args.warmup, 1)
    r = cb.time_wae_cpu(args.batch, N_VOCAB, steps=K, warmup=min(W, 2), budget_s=240.0)
    rc = cb.time_class_cpu(1_000_000)
    sample = '%d of %d requested iterations at batch %d (%.0f ms/step), all %d host threads' % (
        r['steps_timed'], K, args.batch, r['ms_per_step'], r['cores'])
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': r['seq_per_s'], 'unit': UNIT,
        'n_gpus': int(os.environ.get('WORLD_SIZE', '1')), 'steps': K, 'warmup': W, 'ms_per_step': r['ms_per_step'],
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'per_gpu_batch': args.batch, 'seq_len': SEQ_LEN, 'n_vocab': N_VOCAB,
                   'parallelism': 'cpu'},
        'cpu_baseline': {'value': r['seq_per_s'], 'unit': UNIT, 'cores': r['cores'], 'kind': r['kind'], 'sample': sample},
        'e2e': {'value': r['seq_per_s'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
        'class': {'metric': 'class_accepted_samples_per_s', 'value': rc['accepted_per_s'], 'unit': 'samples/s',
                  'draws_per_s': rc['draws_per_s'], 'accept_rate': rc['accept_rate'], 'kind': rc['kind']},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=BATCH)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.gpus > 1 and world == 1:
        # convenience: relaunch under torchrun (the driver launches torchrun itself for N > 1)
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(args.gpus),
               '--master-addr', '127.0.0.1', '--master-port', '29517', os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args)


if __name__ == '__main__':
    main()
