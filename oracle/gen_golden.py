"""Generate the committed golden fixtures under tests/golden/ from the LIVE,
unmodified reference (imported from /root/reference; build container only) and
cross-check the oracle restatement against it while doing so.

    python -m oracle.gen_golden [--out tests/golden] [--skip-b4096]

TEST INFRASTRUCTURE ONLY.  Writes .npz fixtures + REPORT.txt (versions, the
oracle-vs-reference deviations observed at generation time).
"""
import argparse
import copy
import io
import contextlib
import json
import math
import os
import sys
import time
import zlib

import numpy as np
import torch

from . import refharness as rh
from . import wae as ow
from . import class_sampling as oc
from . import decode as od

N_VOCAB = 24
REPORT = []


def say(*a):
    msg = ' '.join(str(x) for x in a)
    print(msg, flush=True)
    REPORT.append(msg)


def digest_indices(name, numel, n=32):
    rs = np.random.RandomState(zlib.crc32(name.encode()) & 0x7fffffff)
    return rs.randint(0, numel, size=min(n, numel))


def digest(name, t):
    """Compact fingerprint of a tensor: fp64 sum, L2 norm, 32 sampled entries."""
    a = t.detach().double().reshape(-1).numpy()
    idx = digest_indices(name, a.size)
    return np.concatenate([[a.sum(), np.sqrt((a * a).sum())], a[idx]])


def state_to_np(sd):
    return {k: v.detach().numpy().copy() for k, v in sd.items()}


def rel_err(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def versions():
    import sklearn
    import scipy
    return json.dumps({'torch': torch.__version__, 'numpy': np.__version__,
                       'sklearn': sklearn.__version__, 'scipy': scipy.__version__,
                       'reference': 'IBM/controlled-peptide-generation@1ba3ce8'})


# ------------------------------------------------------------------ WAE iterations
def run_reference_training(model, tokens_list, noises, lr=1e-3):
    """Drive the reference's own `train_vae.train_vae` (train_vae.py:13-68) for
    len(noises) iterations on injected noise.  Returns per-iteration records."""
    ref = rh.load_reference()
    n_it = len(noises)
    cfgv = ref.cfg.vae          # Bunch (cfg.py:8-11) does not survive deepcopy; edit in place
    cfgv.s_iter, cfgv.n_iter = 0, n_it - 1
    cfgv.cheaplog_every, cfgv.expsvlog_every = 1, 10 ** 9
    cfgv.lr = lr
    rh.reset_rf_cache()
    rh.LOG.values.clear()
    inj = rh.NoiseInjector()
    rec = {'params_after': [], 'grads': [], 'fwd': []}
    vae_names = list(ow.UNIQUE_VAE_PARAMS)
    sd_ref = dict(model.named_parameters())

    def snapshot(i):
        # called at the top of iteration i: params/grads are those left by iteration i-1
        if i > 0:
            rec['params_after'].append({k: sd_ref[k].detach().clone() for k in vae_names})
            rec['grads'].append({k: sd_ref[k].grad.detach().clone() for k in vae_names})
        if i < n_it:
            rh.push_iteration_noise(inj, noises[i], first_iteration=(i == 0))

    ds = rh.DatasetShim(tokens_list, on_batch=snapshot)
    real_forward = model.forward

    def recording_forward(*a, **k):
        out = real_forward(*a, **k)
        (mu, lv), (z, c), logits = out
        rec['fwd'].append({'mu': mu.detach().clone(), 'logvar': lv.detach().clone(),
                           'z': z.detach().clone(), 'c': c.detach().clone(),
                           'logits': logits.detach().clone()})
        return out
    model.forward = recording_forward
    try:
        with inj.active(), contextlib.redirect_stdout(io.StringIO()):
            ref.train_vae.train_vae(cfgv, model, ds)
    finally:
        del model.forward
    snapshot(n_it)
    rec['logged'] = [dict(rh.LOG.values[i]) for i in range(n_it)]
    return rec


LOG_KEYS = ['train_z_mu_L1', 'train_z_logvar', 'train_z_logvar_L1', 'train_z_logvar_KL_penalty',
            'train_L_vae', 'train_L_vae_recon', 'train_L_vae_kl', 'train_L_wae_mmd',
            'train_L_wae_mmdrf', 'train_beta']
ORACLE_KEYS = ['z_mu_l1', 'z_logvar_mean', 'logvar_l1', 'logvar_kl', 'loss', 'recon', 'kl', 'mmd',
               'mmdrf', 'beta']


def beta_for(it):
    ref = rh.load_reference()
    b = ref.cfg.vae.beta
    return ow.anneal_beta(it, b.start.val, b.end.val, b.start.iter, b.end.iter)


def gen_wae(out_dir, batch, n_it, init_sd, full_tensors):
    say('== WAE iterations  B=%d  its=%d' % (batch, n_it))
    model = rh.build_model(N_VOCAB)
    model.load_state_dict(init_sd)
    tokens = [ow.synthetic_tokens(batch, N_VOCAB, seed=1238 + 7 * i) for i in range(n_it)]
    noises = [ow.draw_noise(batch, seed=100 + i) for i in range(n_it)]
    for nz in noises[1:]:
        for k in ('rf_w', 'rf_b', 'rf_u'):
            nz[k] = noises[0][k]
    t0 = time.time()
    rec = run_reference_training(model, tokens, noises)
    say('   reference ran in %.1fs' % (time.time() - t0))

    # oracle on the same inputs
    p = {k: v.clone() for k, v in init_sd.items()}
    state = {}
    worst = 0.0
    o_scal = []
    for i in range(n_it):
        scal, raw_grads, aux = ow.train_step(p, state, tokens[i], noises[i], it=i, beta=beta_for(i))
        o_scal.append(scal)
        for lk, okk in zip(LOG_KEYS, ORACLE_KEYS):
            e = abs(scal[okk] - rec['logged'][i][lk]) / (abs(rec['logged'][i][lk]) + 1e-12)
            worst = max(worst, e)
            if e > 1e-4:
                say('   !! it%d %s oracle %.8g ref %.8g' % (i, okk, scal[okk], rec['logged'][i][lk]))
        for k in ('mu', 'logvar', 'z', 'logits'):
            worst = max(worst, rel_err(aux[k], rec['fwd'][i][k]))
        for k in ow.UNIQUE_VAE_PARAMS:
            e = rel_err(p[k], rec['params_after'][i][k])
            if e > 1e-4:
                say('   !! it%d param %s rel err %.3g' % (i, k, e))
            worst = max(worst, e)
    say('   oracle vs live reference: worst rel deviation %.3g' % worst)

    fx = {'versions': versions(), 'n_vocab': N_VOCAB, 'batch': batch, 'n_it': n_it,
          'token_seeds': np.array([1238 + 7 * i for i in range(n_it)]),
          'noise_seeds': np.array([100 + i for i in range(n_it)]),
          'betas': np.array([beta_for(i) for i in range(n_it)]),
          'logged_keys': np.array(LOG_KEYS),
          'logged': np.array([[rec['logged'][i][k] for k in LOG_KEYS] for i in range(n_it)])}
    for i in range(n_it):
        for k in ow.UNIQUE_VAE_PARAMS:
            fx['it%d/grad_digest/%s' % (i, k)] = digest(k, rec['grads'][i][k])
            fx['it%d/param_digest/%s' % (i, k)] = digest(k, rec['params_after'][i][k])
        fx['it%d/mu_digest' % i] = digest('mu', rec['fwd'][i]['mu'])
        fx['it%d/logvar_digest' % i] = digest('logvar', rec['fwd'][i]['logvar'])
        fx['it%d/logits_digest' % i] = digest('logits', rec['fwd'][i]['logits'])
    if full_tensors:
        for i in range(n_it):
            fx['it%d/tokens' % i] = tokens[i].numpy()
            for k, v in noises[i].items():
                fx['it%d/noise/%s' % (i, k)] = v.numpy()
        for k in ('mu', 'logvar', 'z', 'logits'):
            fx['it0/' + k] = rec['fwd'][0][k].numpy()
        for k in ow.UNIQUE_VAE_PARAMS:
            fx['it0/grad/' + k] = rec['grads'][0][k].numpy()
            fx['final/param/' + k] = rec['params_after'][-1][k].numpy()
    np.savez_compressed(os.path.join(out_dir, 'wae_b%d.npz' % batch), **fx)
    return model


# ------------------------------------------------------------ inference forwards
def gen_infer(out_dir, init_sd):
    say('== encoder / CNN-classifier inference forward (q_c=classifier, sample_z=max)')
    model = rh.build_model(N_VOCAB)
    model.load_state_dict(init_sd)
    model.eval()
    tokens = ow.synthetic_tokens(48, N_VOCAB, seed=4242)
    inj = rh.NoiseInjector()      # WordDropout draws even in eval (decoder.py:117-133): inject "keep all"
    inj.push('binomial', np.zeros(tuple(tokens.shape), dtype='int64'))
    with inj.active(), torch.no_grad():
        (mu, lv), (z, c), logits = model(tokens, q_c='classifier', sample_z='max')
        cnn_logits = model.forward_classifier(tokens)
    p = dict(init_sd)
    o_mu, o_lv = ow.encoder_forward(p, tokens)
    o_cnn = od.cnn_classifier_forward(p, tokens)
    o_logits = ow.decoder_forward(p, tokens, o_mu, torch.softmax(o_cnn, 1), None)
    say('   oracle vs reference: mu %.2g logvar %.2g cnn %.2g dec-logits %.2g' % (
        rel_err(o_mu, mu), rel_err(o_lv, lv), rel_err(o_cnn, cnn_logits), rel_err(o_logits, logits)))
    np.savez_compressed(os.path.join(out_dir, 'infer_b48.npz'), versions=versions(),
                        tokens=tokens.numpy(), mu=mu.numpy(), logvar=lv.numpy(), c=c.numpy(),
                        cnn_logits=cnn_logits.numpy(), dec_logits=logits.numpy())


# ------------------------------------------------------------------------- CLaSS
def synthetic_encodings(n, seed=1238):
    """mu ~ N(0, 0.8^2), logvar ~ N(-2, 0.1^2)."""
    g = torch.Generator().manual_seed(seed)
    mu = 0.8 * torch.randn(n, ow.Z_DIM, generator=g)
    lv = -2.0 + 0.1 * torch.randn(n, ow.Z_DIM, generator=g)
    return mu, lv


def fit_z_classifier(x, seed):
    """sample_pipeline.py:169-192: LogisticRegression(lbfgs, 200) on z-space points
    (synthetic labels from a random hyperplane + noise)."""
    from sklearn.linear_model import LogisticRegression
    rs = np.random.RandomState(seed)
    w = rs.randn(x.shape[1]) / math.sqrt(x.shape[1])
    y = ((x @ w + 0.3 * rs.randn(x.shape[0])) > 0).astype(np.float64)
    clf = LogisticRegression(solver='lbfgs', max_iter=200)
    clf.fit(x, y)
    return clf


def gen_class(out_dir, n_fit=2000, n_comp=100, n_draw=1024):
    say('== CLaSS: mogQ fit + rejection_sample + log densities')
    ref = rh.load_reference()
    dm = ref.density_modeling
    torch.manual_seed(1238)
    np.random.seed(1238)
    mu, lv = synthetic_encodings(n_fit)
    t0 = time.time()
    with contextlib.redirect_stdout(io.StringIO()):
        Q = dm.mogQ(mu, lv, n_components=n_comp, z_num_samples=10, covariance_type='diag')
    say('   GMM fit %.1fs (K=%d on %d x %d)' % (time.time() - t0, n_comp, 10 * n_fit, ow.Z_DIM))
    mun = mu.numpy()
    clfs = {'amp': fit_z_classifier(mun, 11), 'tox': fit_z_classifier(mun, 12)}
    Q.init_attr_classifiers(clfs, clf_targets={'amp': 1, 'tox': 0})   # sample_pipeline.py:290
    np.random.seed(4321)
    z, scores, accepted = Q.rejection_sample(n_samples=n_draw)

    # oracle replay on the same MT19937 stream
    w, m, cv = Q.mog.weights_, Q.mog.means_, Q.mog.covariances_
    rs = np.random.RandomState(4321)
    oz, comp = oc.gmm_sample(w, m, cv, n_draw, rs)
    u = rs.uniform(size=n_draw)
    spec = [('amp', clfs['amp'].coef_[0], clfs['amp'].intercept_[0], 1),
            ('tox', clfs['tox'].coef_[0], clfs['tox'].intercept_[0], 0)]
    oscore, oacc = oc.rejection_accept(oz, u, spec)
    say('   oracle vs reference: z bit-equal %s; accept mask equal %s; accum rel %.2g; rate %.4f' % (
        bool((oz == z.numpy()).all()), bool((oacc == accepted).all()),
        rel_err(oscore['clfZ_prob_accum'], scores['clfZ_prob_accum']), accepted.mean()))

    # densities
    pts = z[:64]
    ref_lq = np.array([Q.logpdf(x) for x in pts])
    ref_lp = np.array([dm.prior_logpdf(x) for x in pts])
    o_lq = oc.gmm_logpdf(pts.numpy(), w, m, cv)
    o_lp = oc.prior_logpdf(pts.numpy())
    say('   logpdf: mog rel %.2g prior rel %.2g' % (rel_err(o_lq, ref_lq), rel_err(o_lp, ref_lp)))
    torch.manual_seed(77)
    nllq, nllp = dm.evaluate_nll(Q, (mu[:64], lv[:64]))
    torch.manual_seed(77)
    sn = np.array([torch.randn(1).item() for _ in range(64)])
    zz = oc.evaluate_nll_points(mun[:64], lv[:64].numpy(), sn)
    o_nllq, o_nllp = -oc.gmm_logpdf(zz, w, m, cv).mean(), -oc.prior_logpdf(zz).mean()
    say('   evaluate_nll: q %.6f/%.6f  p %.6f/%.6f (ref/oracle)' % (nllq, o_nllq, nllp, o_nllp))

    np.savez_compressed(
        os.path.join(out_dir, 'class_sampling.npz'), versions=versions(),
        gmm_weights=w, gmm_means=m, gmm_covs=cv,
        amp_coef=clfs['amp'].coef_[0], amp_b=clfs['amp'].intercept_, tox_coef=clfs['tox'].coef_[0],
        tox_b=clfs['tox'].intercept_, draw_seed=4321, z=z.numpy(), comp=comp, u=u,
        score_amp=scores['clfZ_amp=1'], score_tox=scores['clfZ_tox=0'],
        score_accum=scores['clfZ_prob_accum'], accepted=accepted,
        logpdf_q=ref_lq, logpdf_p=ref_lp, nll_mu=mun[:64], nll_logvar=lv[:64].numpy(),
        nll_noise=sn, nll_q=nllq, nll_p=nllp)


# ------------------------------------------------------------------------ decode
def gen_decode(out_dir, trained_sd, init_sd, mb=64):
    say('== beam / greedy decode of z')
    fx = {'versions': versions()}
    for tag, sd, n in (('trained', trained_sd, mb), ('init', init_sd, 16)):
        model = rh.build_model(N_VOCAB)
        model.load_state_dict(sd)
        g = torch.Generator().manual_seed(31337)
        z = torch.randn(n, ow.Z_DIM, generator=g)
        c_np = np.random.RandomState(5).multinomial(1, [0.5, 0.5], n)
        inj = rh.NoiseInjector()
        inj.push('multinomial', c_np)
        with inj.active(), torch.no_grad():
            hyps, _, c_ix = model.generate_sentences(n, z, sample_mode='beam', beam_size=5)
        c = torch.from_numpy(c_np.astype('float32'))
        with torch.no_grad():
            greedy, _, _ = model.generate_sentences(n, z, c, sample_mode='greedy')
        ref_h = [[[int(t) for t in h] for h in hs] for hs in hyps]
        o_h, margins = od.beam_decode(dict(sd), z, c)
        o_g = od.greedy_decode(dict(sd), z, c)
        same = sum(1 for a, b in zip(ref_h, o_h) if a == b)
        same1 = sum(1 for a, b in zip(ref_h, o_h) if a[0] == b[0])
        say('   [%s] beam: oracle == reference on %d/%d samples (top-1: %d/%d); min margin %.3g; '
            'greedy equal %s' % (tag, same, n, same1, n, min(margins),
                                 bool(o_g.shape == greedy.shape and (o_g == greedy).all())))
        L = ow.MAX_SEQ_LEN + 1
        arr = np.full((n, 3, L), -1, dtype=np.int64)
        for j, hs in enumerate(ref_h):
            for i, h in enumerate(hs):
                arr[j, i, :len(h)] = h
        fx[tag + '/z'] = z.numpy()
        fx[tag + '/c'] = c.numpy()
        fx[tag + '/beam_hyps'] = arr
        fx[tag + '/beam_margin'] = np.array(margins)
        fx[tag + '/greedy'] = greedy.numpy()
    np.savez_compressed(os.path.join(out_dir, 'decode.npz'), **fx)


def train_a_little(init_sd, iters=300, batch=32):
    """A few hundred reference iterations so decode goldens see non-trivial
    weights (sharper logits, learned <eos> placement)."""
    ref = rh.load_reference()
    model = rh.build_model(N_VOCAB)
    model.load_state_dict(init_sd)
    cfgv = ref.cfg.vae
    cfgv.s_iter, cfgv.n_iter = 0, iters
    cfgv.cheaplog_every, cfgv.expsvlog_every = 10 ** 9, 10 ** 9
    rh.reset_rf_cache()
    torch.manual_seed(99)
    np.random.seed(99)
    toks = [ow.synthetic_tokens(batch, N_VOCAB, seed=9000 + i) for i in range(64)]

    class Cyc(rh.DatasetShim):
        def next_batch(self, name):
            self.i += 1
            import types
            return types.SimpleNamespace(text=self.batches[self.i % len(self.batches)])
    with contextlib.redirect_stdout(io.StringIO()):
        ref.train_vae.train_vae(cfgv, model, Cyc(toks))
    rh.reset_rf_cache()
    return {k: v.detach().clone() for k, v in model.state_dict().items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden'))
    ap.add_argument('--skip-b4096', action='store_true')
    args = ap.parse_args()
    out = os.path.abspath(args.out)
    os.makedirs(out, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    say('versions', versions())

    model0 = rh.build_model(N_VOCAB)                      # torch default init under seed 1238
    init_sd = {k: v.detach().clone() for k, v in model0.state_dict().items()}
    np.savez_compressed(os.path.join(out, 'params_init_v24.npz'), **state_to_np(init_sd))

    gen_wae(out, 5, 3, init_sd, full_tensors=True)        # --tiny batch (cfg.py:85-92)
    gen_wae(out, 32, 3, init_sd, full_tensors=True)       # default batch (cfg.py:172)
    if not args.skip_b4096:
        gen_wae(out, 4096, 1, init_sd, full_tensors=False)  # BASELINE config 2; inputs from seeds
    gen_infer(out, init_sd)
    gen_class(out)
    trained_sd = train_a_little(init_sd)
    np.savez_compressed(os.path.join(out, 'params_trained_v24.npz'), **state_to_np(trained_sd))
    gen_decode(out, trained_sd, init_sd)
    with open(os.path.join(out, 'REPORT.txt'), 'w') as fh:
        fh.write('\n'.join(REPORT) + '\n')


if __name__ == '__main__':
    main()
