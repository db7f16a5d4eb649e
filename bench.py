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


# ---- roofline model ---------------------------------------------------------------------------
# ALGORITHMIC work of one launch at per-GPU batch B: the bytes / FLOPs the path needs whatever the implementation
# moves -- h-stash written once by the forward recurrences and read once by BPTT (25*(160+102)*4 B/seq each way),
# tokens, logits only when materialised, mu/logvar/z/eps, parameters + gradients + Adam state once per step; FLOPs of
# the genuine contractions (recurrent, [z;c] projection, heads, fc, RF map, full-kernel MMD Gram).  `frac` in the
# bench line is this figure / event time / measured peak; what the kernel ACTUALLY moves (ncu dram bytes,
# profiles/ncu_traffic.json) is reported next to it as frac_dram and traffic_ratio = dram / algorithmic.
def kernel_work(B, L=SEQ_LEN, V=N_VOCAB, R=500):
    He, Hd, Z, E = 80, 102, 100, 150
    f4 = 4
    rec_enc = 2 * 2 * B * L * He * 3 * He            # both directions
    rec_dec = 2 * B * L * Hd * 3 * Hd
    h_enc, h_dec = 2 * B * L * He * f4, B * L * Hd * f4
    n_par = 258568
    w = {
        'k_gru_fwd_enc': (rec_enc, h_enc + B * L), 'k_gru_fwd_dec': (rec_dec, h_dec + B * L),
        # BPTT: W_hh^T dg contraction + (when fused) the dW_hh / token-table contractions; reads the h stash
        'k_gru_bwd_enc': (2 * rec_enc, h_enc), 'k_gru_bwd_dec': (2 * rec_dec, h_dec),
        'k_wgrad_tc_enc': (rec_enc // 2, 0), 'k_wgrad_tc_dec': (rec_dec, 0),     # part of BPTT algorithmically: no bytes of their own
        'k_wgrad_hh_enc': (rec_enc // 2, 0), 'k_wgrad_hh_dec': (rec_dec, 0),
        'k_dec_out': (2 * 3 * B * L * Hd * V, 0),      # fc fwd + 2 bwd products; h already counted with the recurrences
        'k_mmd_gram': (3 * 2 * B * B * Z, 2 * B * Z * f4),
        'k_clip_adam': (0, 7 * n_par * f4), 'k_sumsq_partial': (0, n_par * f4),
        'k_prep_tokens': (0, B * L * 8), 'k_step_noise': (0, B * (3 * Z * f4 + L * Hd + L + 8)),
        'k_input_grads': (2 * V * E * (2 * 3 * He + 3 * Hd) * 2, 0),
        'k_prep_weights': (2 * V * E * (2 * 3 * He + 3 * Hd), 2 * n_par * f4),
        # dense layers around the latent code (heads, [z;c] projection; their backward + the head weight gradients):
        # mu / logvar / z written + eps read, resp. mu / logvar / eps read
        'k_latent_fwd_tc': (2 * B * (2 * Z * 2 * He + 3 * Hd * (Z + 2)), 4 * B * Z * f4),
        'k_latent_bwd_tc': (2 * B * (3 * Hd * (Z + 2) + 2 * 2 * Z * (2 * He)), 3 * B * Z * f4),
        # random-feature map (z read) and its gradient (dz written)
        'k_rf_feat_tc': (2 * B * Z * R, B * Z * f4), 'k_rf_grad_tc': (2 * B * Z * R, B * Z * f4),
        # dW_ih[:,150:] and the head weight / bias gradients (two launches of one kernel: the mean of the two shapes)
        'k_wgrad_dense_tc': (B * (3 * Hd * (Z + 2) + 2 * Z * (2 * He + 1)), 0),
    }
    # the dense layers, by shape label "MxNxK": heads (B x 100 x 160, x2 + transposes), [z;c] projection, RF map
    return w


def work_for(label, work, B):
    """(flops, bytes) of one launch of kernel `label`; dense products carry their shape in the label."""
    import re
    m = re.match(r'k_(?:sgemm|gemm_tc)\w*\[(\d+)x(\d+)x(\d+)(?:x(\d+))?\]', label)
    if m:
        M, N, K = int(m.group(1)), int(m.group(2)), int(m.group(3))
        n = int(m.group(4) or 1)
        return 2 * M * N * K * n, 0                 # operands are activations already counted (mu/z/h) or parameters
    for k in sorted(work, key=len, reverse=True):
        if label.startswith(k):
            return work[k]
    return None


def step_work(B, L=SEQ_LEN, V=N_VOCAB):
    """60 KB/seq + params/grads/Adam once; 11.8 MFLOP/seq + the MMD Gram."""
    seq_bytes = 200 + 2 * L * (160 + 102) * 4 + 2 * L * V * 4 + 3200
    return {'alg_bytes': B * seq_bytes + 6.2e6, 'alg_flops': B * 11.8e6 + 3 * 2 * B * B * 100}


def build_roofline(rows, K, B, ms_per_step, peaks, peak_src):
    """rows: (label, total ms, launches) over K serialised steps -> the `roofline` object of the bench line."""
    work = kernel_work(B)
    try:
        with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
            traffic = json.load(f)
    except (OSError, ValueError):
        traffic = {}
    hbm, tens = peaks['hbm_gbs'] * 1e9, peaks['bf16_tflops_sustained'] * 1e12
    tot_ms = sum(ms for _, ms, _ in rows) / K
    per, dram_known = {}, 0.0
    for label, ms, cnt in rows:
        kms = ms / max(cnt, 1)
        wk = work_for(label, work, B)
        tr = traffic.get(label)
        ent = {'ms': round(kms, 4), 'launches_per_step': cnt / K, 'share_of_step': round(ms / K / tot_ms, 4)}
        if wk is not None and kms > 0:
            fl, by = wk
            ent.update({'alg_flops': fl, 'alg_bytes': by,
                        'frac_hbm_algorithmic': round(by / (kms / 1e3) / hbm, 4),
                        'frac_tensor_algorithmic': round(fl / (kms / 1e3) / tens, 4)})
        if tr is not None and kms > 0:
            ent.update({'traffic': tr, 'frac_dram': round(tr / (kms / 1e3) / hbm, 4)})
            if wk is not None and wk[1] > 0:
                ent['traffic_ratio'] = round(tr / wk[1], 2)
            dram_known += tr * cnt / K
        per[label] = ent
    dom = max(per, key=lambda k: per[k]['share_of_step'] )
    d = per[dom]
    fh, ft = d.get('frac_hbm_algorithmic', 0.0), d.get('frac_tensor_algorithmic', 0.0)
    kms = d['ms'] / 1e3
    if ft > fh:
        roof = {'bound': 'tensor', 'achieved': round(d.get('alg_flops', 0) / kms / 1e12, 3), 'peak': peaks['bf16_tflops_sustained'],
                'unit': 'TFLOP/s', 'frac': ft}
    else:
        roof = {'bound': 'hbm', 'achieved': round(d.get('alg_bytes', 0) / kms / 1e9, 1), 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                'frac': fh}
    sw = step_work(B)
    step = {'ms': round(ms_per_step, 4), 'alg_bytes': sw['alg_bytes'], 'alg_flops': sw['alg_flops'],
            'frac_hbm_algorithmic': round(sw['alg_bytes'] / (ms_per_step / 1e3) / hbm, 4),
            'frac_tensor_algorithmic': round(sw['alg_flops'] / (ms_per_step / 1e3) / tens, 4),
            'dram_bytes_captured_kernels': dram_known,
            'frac_hbm_dram': round(dram_known / (ms_per_step / 1e3) / hbm, 4) if dram_known else None,
            'traffic_ratio': round(dram_known / sw['alg_bytes'], 2) if dram_known else None}
    roof.update({'kernel': dom, 'kernel_ms': d['ms'], 'launches_per_step': d['launches_per_step'],
                 'share_of_step': d['share_of_step'], 'frac_algorithmic': roof['frac'], 'frac_dram': d.get('frac_dram'),
                 'traffic': d.get('traffic'), 'traffic_ratio': d.get('traffic_ratio'), 'step': step,
                 'per_kernel': dict(sorted(per.items(), key=lambda kv: -kv[1]['share_of_step'])),
                 'peak_source': peak_src + (' (MEASURED_PEAKS.json: HBM copy GB/s, bf16 sustained TFLOP/s)' if peak_src == 'measured' else ''),
                 'note': 'dominant kernel = largest share of the step by CUDA events (serialised pass); frac = ALGORITHMIC bytes or '
                         'FLOPs of that kernel (SURVEY 8d: h-stash once each way, genuine contractions) / event time / measured peak; '
                         'frac_dram = ncu dram bytes of one launch (profiles/ncu_traffic.json) / event time / HBM peak; '
                         'traffic_ratio = dram / algorithmic bytes; step = whole iteration against 60 KB/seq + 11.8 MFLOP/seq'})
    return roof


def setup_model(device):
    import numpy as np
    import torch
    import cfg
    from models.model import RNN_VAE
    torch.manual_seed(cfg.seed)
    np.random.seed(cfg.seed)
    model = RNN_VAE(n_vocab=N_VOCAB, max_seq_len=cfg.max_seq_len, **cfg.model).to(device)
    return cfg, model


def run_strong(args, cfg, model, st, dev, world, rank, timed, hp):
    """Global batch 4096 split over the ranks (512 per GPU at N = 8); the log-only full-kernel MMD is the GLOBAL-batch
    value (all-gather of z).  Both kernel families are timed: the fp32 SIMT kernels the library picks by
    itself below 512 rows per GPU, and the tcgen05 kernels forced on."""
    import torch
    import utils
    from cpg_b200 import _lib, engine, parallel, synth
    Bs = max(1, BATCH // world)
    tok = synth.synthetic_tokens(Bs, N_VOCAB, seed=4242 + rank).to(dev)
    noise = engine.alloc_noise(Bs, SEQ_LEN, dev, seed=cfg.seed)
    seed = cfg.b200.noise_seed + 7919 * rank
    out = {'global_batch': Bs * world, 'per_gpu_batch': Bs, 'full_mmd': 'global', 'steps': args.steps}
    it = {'n': 0}

    def step():
        it['n'] += 1
        hp.beta = float(utils.anneal(cfg.vae.beta, it['n']))
        engine.fill_step_noise(noise, seed, it['n'], overlap=True)
        return parallel.dp_train_step(st, tok, noise, hp, global_batch=Bs * world, full_mmd='global')
    try:
        for name, flag in (('auto', 1), ('tcgen05_forced', 2)):
            for o in ('gru_tensor_core', 'dec_out_tensor_core', 'wgrad_tensor_core'):
                _lib.set_option(o, flag)
            for _ in range(3):
                step()
            ms = timed(step, args.steps) / args.steps
            out[name] = {'ms_per_step': ms, 'seq_per_s': Bs * world / (ms / 1e3)}
    finally:
        for o in ('gru_tensor_core', 'dec_out_tensor_core', 'wgrad_tensor_core'):
            _lib.set_option(o, 1)
    best = max(('auto', 'tcgen05_forced'), key=lambda k: out[k]['seq_per_s'])
    out.update({'value': out[best]['seq_per_s'], 'unit': UNIT, 'kernels': best, 'scaling': 'strong'})
    return out


def run_class(args, st, dev, world, rank, barrier, model):
    """CLaSS sampling, sharded by Philox offset: rank r draws n per-GPU samples with global indices [r n, (r+1) n); no
    data-path collective, the accepted counts are summed at the end of the round.  Three measurements:
      value     device-resident rejection sampling with the reference's full return contract in HBM (z fp32, 3 score
                arrays fp64, mask: 425 B/draw), all ranks, max time over ranks
      e2e       through density_modeling.mogQ.rejection_sample(n): the same plus the copies to host tensors it returns
      pipeline  one device round of config 5: flags only -> compaction -> re-generation of accepted z -> beam decode ->
                dedup -> descriptors -> unique accepted peptides to the host (mogQ.rejection_sample_decode)."""
    import numpy as np
    import sklearn.mixture
    import torch
    import torch.distributed as dist
    import density_modeling as dm
    from cpg_b200 import sampling, synth
    w, means, covs, clfs = synth.synthetic_class_setup()
    gmm = sampling.GmmDevice(w, means, covs, dev)
    spec = sampling.ClassifierSpec(clfs, dev)
    n = args.class_draws or (10_000_000 if world == 1 else 12_500_000)
    peaks, _ = load_peaks()

    def reduce_max(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t)
        return float(t.item())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # (1) device-resident, full contract
    sampling.class_sample(gmm, spec, n, 1, offset=rank * n)
    barrier()
    e0.record()
    out = sampling.class_sample(gmm, spec, n, 2, offset=rank * n)
    e1.record()
    torch.cuda.synchronize()
    ms_local = e0.elapsed_time(e1)
    cms = reduce_max(ms_local)
    acc = reduce_sum(float(out['n_accepted'].item()))
    del out
    # flags only (what the device pipeline runs)
    sampling.class_sample(gmm, spec, n, 3, offset=rank * n, want_z=False, want_scores=False)
    barrier()
    e0.record()
    sampling.class_sample(gmm, spec, n, 4, offset=rank * n, want_z=False, want_scores=False)
    e1.record()
    torch.cuda.synchronize()
    fms = reduce_max(e0.elapsed_time(e1))
    torch.cuda.empty_cache()
    # (2) end to end through the reference-facing call (host outputs), bounded sample
    mog = sklearn.mixture.GaussianMixture(n_components=len(w), covariance_type='diag')
    mog.weights_, mog.means_, mog.covariances_ = w, means, covs
    mog.precisions_cholesky_ = 1.0 / np.sqrt(covs)
    Q = dm.mogQ.from_fitted(mog)
    mk = lambda c: types.SimpleNamespace(coef_=np.asarray(c[1])[None, :], intercept_=np.asarray([c[2]]))
    Q.init_attr_classifiers({c[0]: mk(c) for c in clfs}, {c[0]: c[3] for c in clfs})
    Q._draw_offset = rank * (1 << 40)
    n_e2e = min(n, 2_000_000)
    Q.rejection_sample(100_000)
    barrier()
    t0 = time.perf_counter()
    z_h, scores_h, accepted = Q.rejection_sample(n_e2e)
    d2h = (z_h.numel() * z_h.element_size() + sum(v.nbytes for v in scores_h.values()) + accepted.nbytes) / n_e2e
    torch.cuda.synchronize()
    e2e_s = reduce_max(time.perf_counter() - t0)
    e2e_acc = reduce_sum(float(accepted.sum()))
    # (3) device round with decode of the accepted z
    aa = ['<unk>', '<pad>', '<start>', '<eos>'] + list('ACDEFGHIKLMNPQRSTVWY')
    ds = types.SimpleNamespace(idx2sentences=lambda seqs, print_special_tokens=True: [
        ' '.join(aa[int(i)] for i in s_ if print_special_tokens or int(i) > 3) for s_ in seqs])
    n_pipe = min(n, 2_000_000)
    Q.rejection_sample_decode(100_000, model, ds)
    barrier()
    t0 = time.perf_counter()
    df = Q.rejection_sample_decode(n_pipe, model, ds)
    torch.cuda.synchronize()
    pipe_s = reduce_max(time.perf_counter() - t0)
    stats = Q.last_round_stats
    pipe_acc, pipe_unique = reduce_sum(stats['n_accepted']), reduce_sum(stats['n_unique'])
    # device-only timing of the same round (no host tables)
    barrier()
    e0.record()
    Q.rejection_sample_decode(n_pipe, model, ds, return_device=True)
    e1.record()
    torch.cuda.synchronize()
    pipe_dev_ms = reduce_max(e0.elapsed_time(e1))
    # beam decode alone
    zb = torch.randn(8192, 100, device=dev)
    cbm = torch.eye(2, device=dev)[torch.arange(8192, device=dev) % 2]
    sampling.beam_decode(st.params, N_VOCAB, zb, cbm)
    barrier()
    e0.record()
    sampling.beam_decode(st.params, N_VOCAB, zb, cbm)
    e1.record()
    torch.cuda.synchronize()
    bms = reduce_max(e0.elapsed_time(e1))
    tot = n * world
    gbs_rank = 425 * n / (ms_local / 1e3) / 1e9
    # beam decode against the fp32 FMA peak: per beam and step the recurrent gates (3*102 x 102), the output layer (V x 102)
    # and the gate math; 5 beams x 25 steps per decoded z (upper bound: finished samples stop early)
    beam_flops = 2.0 * (3 * 102 * 102 + N_VOCAB * 102) * 5 * 25
    fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12
    beam_tf = beam_flops * 8192 / (bms / 1e3) / 1e12
    beam_roof = {'kernel': 'k_decode<beam>', 'bound': 'fp32 FMA (SIMT: 25 dependent steps of 30-row matrix-vector products per '
                 'CTA, no tensor-core shape)', 'achieved': round(beam_tf, 2), 'peak': round(fp32_peak, 1), 'unit': 'TFLOP/s',
                 'frac': round(beam_tf / fp32_peak, 4), 'flops_per_seq': beam_flops,
                 'peak_source': 'nominal: 148 SMs x 128 FMA lanes x 2 x 1.965 GHz (not in MEASURED_PEAKS.json)'}
    # mogQ.logpdf / evaluate_nll (density_modeling.py:64-73,100-106): diag-GMM log-density in fp64, one pass over the points
    logpdf = None
    try:
        npts = 4_000_000
        xp = torch.randn(npts, 100, device=dev)
        sampling.gmm_logpdf(gmm, xp)
        barrier()
        e0.record()
        lp = sampling.gmm_logpdf(gmm, xp)
        e1.record()
        torch.cuda.synchronize()
        lms = reduce_max(e0.elapsed_time(e1))
        K = int(gmm.K)
        lgb = npts * 408 / (lms / 1e3) / 1e9
        ltf = npts * K * 100 * 3 / (lms / 1e3) / 1e12
        fp64_peak = 40.0
        logpdf = {'kernel': 'k_gmm_logpdf', 'points_per_s': npts * world / (lms / 1e3), 'n_points': npts, 'components': K,
                  'bound': 'fp64 FMA (%d components x 100 dimensions per point: 3 fp64 FLOP per (point, component, dimension) in '
                           'sklearn\'s ordering, results pinned at 1e-9)' % K,
                  'achieved': round(ltf, 2), 'peak': fp64_peak, 'unit': 'TFLOP/s', 'frac': round(ltf / fp64_peak, 4),
                  'peak_source': 'nominal B200 fp64 vector rate (not in MEASURED_PEAKS.json)',
                  'hbm_gbs': round(lgb, 1), 'hbm_frac': round(lgb / peaks['hbm_gbs'], 4),
                  'note': 'a diagnostic (mogQ.logpdf / evaluate_nll), not on the sampling path; 408 algorithmic bytes per point'}
        del xp, lp
    except Exception as exc:                               # a side measurement must never take the line down
        logpdf = {'error': repr(exc)[:200]}
    return {'metric': 'class_accepted_samples_per_s', 'value': acc / (cms / 1e3), 'unit': 'samples/s', 'n_gpus': world,
            'scaling': 'weak' if world > 1 else None, 'sharding': 'Philox offset = rank * n_draws_per_gpu, no collective on the data path',
            'draws_per_s': tot / (cms / 1e3), 'accept_rate': acc / tot, 'n_draws': tot, 'n_draws_per_gpu': n,
            'flags_only_draws_per_s': tot / (fms / 1e3),
            'roofline': {'kernel': 'k_class_sample', 'bound': 'hbm', 'achieved': round(gbs_rank, 1), 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                         'frac': round(gbs_rank / peaks['hbm_gbs'], 4), 'bytes_per_draw': 425,
                         'note': 'algorithmic bytes = the reference return contract (z fp32 + 3 fp64 score arrays + mask) written once'},
            'e2e': {'value': e2e_acc / e2e_s, 'unit': 'samples/s', 'draws_per_s': n_e2e * world / e2e_s, 'n_draws': n_e2e * world,
                    'd2h_bytes_per_draw': d2h, 'h2d_bytes_per_draw': 0,
                    'api': 'density_modeling.mogQ.rejection_sample(n): z, 3 score arrays and the mask returned as host arrays'},
            'pipeline': {'accepted_decoded_per_s': pipe_acc / pipe_s, 'unique_peptides_per_s': pipe_unique / pipe_s,
                         'draws_per_s': n_pipe * world / pipe_s, 'n_draws': n_pipe * world, 'n_accepted': pipe_acc,
                         'n_unique': pipe_unique, 'device_only_accepted_decoded_per_s': pipe_acc / (pipe_dev_ms / 1e3),
                         'api': 'mogQ.rejection_sample_decode(n, model, dataset): flags -> compaction -> re-generation -> beam decode '
                                '-> dedup -> H/uH/charge on the device, unique accepted peptides to a host table'},
            'beam_decode_seq_per_s': 8192 * world / (bms / 1e3), 'beam_roofline': beam_roof, 'logpdf': logpdf}


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
    import utils
    from cpg_b200 import synth

    cfg, model = setup_model(dev)
    st = model.bind_grads()
    B, L, K, W = args.batch, SEQ_LEN, args.steps, args.warmup
    tokens_host = synth.synthetic_tokens(B, N_VOCAB, seed=1238 + rank).pin_memory()
    tokens = tokens_host.to(dev)
    noise = engine.alloc_noise(B, L, dev, seed=cfg.seed)
    hp = engine.make_hparams(lr=cfg.vae.lr, z_regu=cfg.vae.z_regu_loss, mmd_sigma=cfg.losses.wae_mmd.sigma,
                             rf_dim=cfg.losses.wae_mmd.rf_dim)
    seed = cfg.b200.noise_seed + 7919 * rank
    gb = B * world
    counter = {'it': 0}

    stepper = engine.FusedStepper(st, B, L, hp, seed=seed, rf_dim=cfg.losses.wae_mmd.rf_dim) if world == 1 else None
    dp_stepper = parallel.GraphedDPStepper(st, B, L, hp, noise, seed, gb, graph=not args.no_dp_graph) if world > 1 else None

    def step():
        it = counter['it']
        counter['it'] += 1
        beta = float(utils.anneal(cfg.vae.beta, it))
        if world == 1:                                             # what train_vae issues per iteration: noise + step, one C call
            return stepper.step(tokens, it, beta)                  # (captured CUDA graph from the third call on)
        return dp_stepper.step(tokens, it, beta)                   # noise + dp_train_step (collectives included), one captured graph

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

    # ---- reduced-precision mode (NOT the headline): BASELINE.json configs[2] names "bf16 matmul tiles"; option matmul_terms=1
    # keeps only the leading bf16 product of every recurrence / decoder-output contraction (outside the parity bars)
    reduced = None
    if world == 1:
        _lib.set_option('matmul_terms', 1)
        for _ in range(max(W, 3)):
            step()
        r_ms = timed(step, K) / K
        _lib.set_option('matmul_terms', 3)
        for _ in range(3):
            step()
        reduced = {'option': 'matmul_terms=1 (single bf16 product in the recurrences and the decoder-output layer)', 'ms_per_step': r_ms,
                   'seq_per_s': gb / (r_ms / 1e3), 'note': 'reduced precision, outside the parity bars (logits ~2e-3, gradients ~4e-3 of max '
                   'vs the three-product configuration); the headline value above is the three-product fp32-grade configuration'}

    # ---- per-kernel durations (CUDA events around every launch of the same step, same stream)
    # The step overlaps its latency-bound loss / weight-gradient kernels with the recurrences on a side stream;
    # for the per-kernel pass everything is serialised on one stream so that each duration is the kernel's own.
    _lib.set_option('side_stream', 0)
    _lib.profile_enable(True)
    if dp_stepper is not None:
        dp_stepper.force_eager = True                              # the profiler hooks the library's eager launches
    prof_ms = timed(step, K)
    rows = _lib.profile_read()
    _lib.profile_enable(False)
    if dp_stepper is not None:
        dp_stepper.force_eager = False
    _lib.set_option('side_stream', 1)
    peaks, peak_src = load_peaks()
    roof = build_roofline(rows, K, B, ms_per_step, peaks, peak_src)

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

    dp_graph_state = None
    if dp_stepper is not None:
        dp_graph_state = 'captured' if dp_stepper.graph is not None else 'capture unavailable, eager launches'
        dp_stepper.release()                                       # no live graph with captured collectives at teardown

    # ---- strong scaling (BASELINE.json configs[2]: global batch 4096 over the N GPUs), N > 1 only
    strong = None
    if world > 1 and not args.no_strong:
        strong = run_strong(args, cfg, model, st, dev, world, rank, timed, hp)

    # ---- CLaSS (second metric of BASELINE.json: accepted samples/s; configs[3] at N = 1, configs[4] sharded)
    class_block = run_class(args, st, dev, world, rank, barrier, model)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- CPU baseline on this box's host cores (bounded sample)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_baseline as cb          # the ONLY use of oracle/ by this arm: the CPU baseline leg
        r = cb.time_wae_cpu(B, N_VOCAB, steps=5, warmup=1, budget_s=90.0)      # ~13 s of CPU work on 16 host threads
        cpu = {'value': r['seq_per_s'], 'unit': UNIT, 'cores': r['cores'], 'kind': r['kind'],
               'sample': '%d full iterations at batch %d after 1 warm-up (%.0f ms/step)' % (
                   r['steps_timed'], B, r['ms_per_step'])}
        rc = cb.time_class_cpu(1_000_000)
        class_block['cpu_baseline'] = {'value': rc['accepted_per_s'], 'unit': 'samples/s', 'cores': 1, 'kind': rc['kind'],
                                       'sample': 'rejection_sample(1e6): %.2f s, accept rate %.3f' % (rc['seconds'], rc['accept_rate'])}
        rb = cb.time_beam_cpu({k: v.detach().cpu() for k, v in st.views(st.params).items()}, 192)
        class_block['beam_cpu_baseline'] = {'value': rb['seq_per_s'], 'unit': 'seq/s', 'cores': rb['cores'], 'kind': rb['kind'],
                                            'sample': 'beam decode (beam 5, n_best 3) of %d z: %.2f s' % (rb['n'], rb['seconds'])}
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': max(W, 3),
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'fp32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'per_gpu_batch': B, 'global_batch': gb, 'seq_len': L, 'n_vocab': N_VOCAB,
                   'parallelism': 'dp%d' % world if world > 1 else 'single',
                   'l2_policy': 'per-step working set (activation stash ~1.1 GB) exceeds the 126 MB L2; no flush needed',
                   'noise': 'Philox in-kernel, regenerated every step',
                   'launch': ('single GPU: one captured CUDA graph per iteration (cpg_wae_train_step_philox)' if world == 1 else
                              'one captured CUDA graph per rank and iteration, NCCL all-reduces included (parallel.GraphedDPStepper): %s'
                              % dp_graph_state),
                   'arithmetic': 'fp32 storage and accumulation; recurrence / decoder-output contractions as split-bf16 '
                                 '(x1+x2, 3 products; logits 3 terms) tcgen05 MMAs; heads / [z;c] projection / RF map as split-fp16, their '
                                 'backward and the dense weight gradients (dW_ih[:,150:], heads) as split-bf16 tcgen05 MMAs (3 products); '
                                 'MMD Gram tf32'},
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': B * L * 8, 'd2h_bytes_per_step': 16 * 4 + 8,
                'api': 'train_vae.train_vae(cfgv, model, dataset): pinned host tokens copied H2D every step (one step ahead, copy stream), '
                  'scalar block copied D2H every step (collected after the next step is enqueued)'},
        'gpu_launches': launches, 'launches_per_step': launches / K,
        'profiled_ms_per_step': prof_ms / K,
        'roofline': roof, 'cpu_baseline': cpu, 'clocks': clock_summary, 'class': class_block, 'strong_scaling': strong,
        'reduced_precision': reduced,
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
    ap.add_argument('--no-strong', action='store_true', help='skip the strong-scaling arm at N > 1')
    ap.add_argument('--no-dp-graph', action='store_true', help='N > 1: eager launches instead of the captured data-parallel graph')
    ap.add_argument('--class-draws', type=int, default=0, help='CLaSS draws per GPU (default 10M at N=1, 12.5M at N>1)')
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
