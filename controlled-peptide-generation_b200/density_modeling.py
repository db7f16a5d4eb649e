"""Latent density Q(z) + attribute-conditioned rejection sampling with the reference's names
(density_modeling.py of IBM/controlled-peptide-generation).

`mogQ` still FITS with scikit-learn (setup, not hot: reference density_modeling.py:64-73), but
sampling, classifier scoring, the accept test and the log densities run on the GPU:
  rejection_sample(n)            -> cpg_class_sample   (Philox draws; `mode='reference_rng'` replays the
                                    numpy/sklearn stream on the host and scores on the GPU instead)
  logpdf / logpdf_batch          -> cpg_gmm_logpdf
  prior_logpdf / evaluate_nll    -> cpg_prior_logpdf / batched kernels (no per-point Python loop)
"""
import math

import numpy as np
import torch

from cpg_b200 import _lib, sampling


def _device(device=None):
    return _lib.tensor_device(device)


def prior_logpdf(z):
    """log N(z; 0, I) for one point (reference density_modeling.py:11-14)."""
    dev = _device()
    return float(sampling.prior_logpdf(z.reshape(1, -1).float().to(dev))[0])


_RING = {}                                # device index -> (pinned uint8 [2, RING_BYTES], [event, event])
RING_BYTES = 64 << 20


def _to_host_chunked(tensors, dev):
    """Fresh host tensors with the contents of the device tensors (any dtype / shape, contiguous)."""
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _RING:
        _RING[key] = (torch.empty(2, RING_BYTES, dtype=torch.uint8, pin_memory=True), [None, None])
    ring, events = _RING[key]
    stream = torch.cuda.current_stream(dev)
    outs, jobs = [], []
    for t in tensors:
        t = t.contiguous()
        out = torch.empty(t.shape, dtype=t.dtype)
        outs.append(out)
        src, dst = t.view(-1).view(torch.uint8), out.view(-1).view(torch.uint8)
        for lo in range(0, src.numel(), RING_BYTES):
            jobs.append((src, dst, lo, min(lo + RING_BYTES, src.numel())))
    pending = None                                       # (slot, dst, lo, hi) whose D2H copy is in flight
    for k, (src, dst, lo, hi) in enumerate(jobs):
        slot = k & 1
        ring[slot, :hi - lo].copy_(src[lo:hi], non_blocking=True)
        events[slot] = stream.record_event()
        if pending is not None:
            ps, pd, plo, phi = pending
            events[ps].synchronize()
            pd[plo:phi].copy_(ring[ps, :phi - plo])
        pending = (slot, dst, lo, hi)
    if pending is not None:
        ps, pd, plo, phi = pending
        events[ps].synchronize()
        pd[plo:phi].copy_(ring[ps, :phi - plo])
    return outs


class RejSampleBase:
    seed = 1238
    _draw_offset = 0

    def init_attr_classifiers(self, attr_clfs, clf_targets):
        """attr_clfs: ordered dict name -> fitted sklearn LogisticRegression (binary);
        clf_targets: name -> column of predict_proba to keep (sample_pipeline.py:290)."""
        self.attr_clfs = attr_clfs
        self.clf_targets = clf_targets
        self._spec = None

    def _clf_list(self):
        return [(name, np.asarray(clf.coef_).reshape(-1), np.asarray(clf.intercept_).reshape(-1),
                 int(self.clf_targets[name])) for name, clf in self.attr_clfs.items()]

    def _clf_spec(self, dev):
        if getattr(self, '_spec', None) is None or self._spec_dev != dev:
            self._spec = sampling.ClassifierSpec(self._clf_list(), dev)
            self._spec_dev = dev
        return self._spec

    def score_clf(self, attr_name, z):
        """P(attr == target | z) from the z-space classifier (reference density_modeling.py:43-48)."""
        dev = _device()
        name, coef, b, col = [c for c in self._clf_list() if c[0] == attr_name][0]
        spec = sampling.ClassifierSpec([(name, coef, b, col)], dev)
        zd = z.float().to(dev)
        probs, _, _ = sampling.score_accept(zd, torch.zeros(zd.shape[0], dtype=torch.float64, device=dev), spec)
        out = probs[0].cpu().numpy()
        return out.astype(np.float32) if spec.all_f32 else out

    def rejection_sample(self, n_samples, prefix='clfZ', mode='philox', device=None):
        """-> (samples_z float32 [n,100] (CPU tensor), scores dict, accepted bool[n]) like the reference
        (density_modeling.py:50-60).  mode='philox' (default) draws on the GPU; mode='reference_rng'
        consumes numpy's global stream exactly as the reference does (z via sklearn, then the uniforms)
        and only scores / accepts on the GPU -- bit-comparable with the reference under the same seed."""
        dev = _device(device)
        spec = self._clf_spec(dev)
        if mode == 'reference_rng':
            z = self.sample(n_samples).to(dev)
            u = torch.from_numpy(np.random.uniform(size=n_samples)).to(dev)
            probs, accum, accept = sampling.score_accept(z, u, spec)
        else:
            out = sampling.class_sample(self._gmm_device(dev), spec, n_samples, self.seed, self._draw_offset)
            self._draw_offset += n_samples
            z, probs, accum, accept = out['z'], out['probs'], out['accum'], out['accept']
        # device -> host: the reference's return contract is host data -- z as a CPU tensor, scores / mask as numpy arrays.
        # Chunks go through a small pinned ring that lives across calls (pinning hundreds of MB per call costs more than the
        # copy itself); the host copies chunk k out of the ring into the result while chunk k + 1 crosses PCIe.
        if spec.all_f32:                                 # float32 classifiers score in float32 (see the class docstring)
            probs, accum = probs.float(), accum.float()
        hz, hp, ha, hm = _to_host_chunked([z, probs, accum, accept], dev)
        scores_z = {}
        for i, name in enumerate(spec.names):
            scores_z['{}_{}={}'.format(prefix, name, spec.target[i])] = hp[i].numpy()
        scores_z[prefix + '_prob_accum'] = ha.numpy()
        return hz, scores_z, hm.numpy().astype(bool)

    def rejection_sample_decode(self, n_samples, model, dataset, prefix='clfZ', device=None, n_best=3, return_device=False):
        """One sampling round with everything after the draw kept on the GPU (BASELINE.json config 5: "beam decode of
        accepted z"): Philox draws + scores + accept (no z in HBM for rejected draws) -> stream compaction of the
        accept mask -> re-generation of the accepted z -> beam decode -> duplicate removal -> H / uH / charge.
        Only the unique accepted peptides cross PCIe.  Returns a DataFrame with one row per unique accepted peptide
        (columns of the reference's round table + draw_index, H, uH, charge) and sets self.last_round_stats."""
        import pandas as pd
        from cpg_b200 import peptides
        dev = _device(device)
        spec = self._clf_spec(dev)
        gmm = self._gmm_device(dev)
        off = self._draw_offset
        out = sampling.class_sample(gmm, spec, n_samples, self.seed, off, want_z=False, want_scores=False)
        self._draw_offset += n_samples
        n_acc = int(out['n_accepted'].item())                       # the one host sync of the round (8 bytes)
        idx, _ = sampling.compact_accepted(out['accept'], first_index=off, cap=max(n_acc, 1))
        idx = idx[:n_acc]
        z, probs, accum = sampling.class_regen(gmm, spec, self.seed, idx, want_scores=True)
        st = model.flat_state()
        c = model.sample_c_prior(n_acc) if n_acc else torch.zeros(0, 2, device=dev)
        stats = {'n_draws': n_samples, 'n_accepted': n_acc, 'n_unique': 0}
        if n_acc == 0:
            self.last_round_stats = stats
            return pd.DataFrame(columns=['peptide', 'z', 'accept_z', 'draw_index', 'H', 'uH', 'charge'])
        toks, lens, _ = sampling.beam_decode(st.params, model.n_vocab, z, c.to(dev), model.MAX_SEQ_LEN, 5, n_best)
        hyp0 = toks[:, 0, :].contiguous()
        _, is_first = peptides.dedup_rows(hyp0)
        keep = torch.nonzero(is_first, as_tuple=False).squeeze(1)
        words = peptides.vocabulary_words(dataset, model.n_vocab)
        a2t = peptides.token_to_residue(lambda t: words[t], model.n_vocab)
        H, uH, ch, _ = peptides.descriptors_from_tokens(hyp0, a2t)
        stats['n_unique'] = int(keep.numel())
        self.last_round_stats = stats
        if return_device:
            return {'tokens': hyp0, 'keep': keep, 'z': z, 'idx': idx, 'H': H, 'uH': uH, 'charge': ch, 'accum': accum}
        # device -> host: the unique accepted rows only
        # (the table of peptide strings is built without a Python loop over tokens: a round holds ~10^5..10^6 rows)
        seqs = peptides.rows_to_sentences(hyp0[keep].cpu().numpy(), lens[keep, 0].cpu().numpy(), words,
                                          fallback=lambda rows: dataset.idx2sentences(rows, print_special_tokens=False))
        cast = (lambda a: a.astype(np.float32)) if spec.all_f32 else (lambda a: a)
        df = {'peptide': seqs, 'z': list(z[keep].cpu().numpy()), 'accept_z': np.ones(len(seqs), dtype=bool),
              'draw_index': idx[keep].cpu().numpy()}
        for i, name in enumerate(spec.names):
            df['{}_{}={}'.format(prefix, name, spec.target[i])] = cast(probs[i][keep].cpu().numpy())
        df[prefix + '_prob_accum'] = cast(accum[keep].cpu().numpy())
        df['H'], df['uH'], df['charge'] = H[keep].cpu().numpy(), uH[keep].cpu().numpy(), ch[keep].cpu().numpy()
        return pd.DataFrame(df)


class mogQ(RejSampleBase):
    def __init__(self, mu, logvar, n_components=10, z_num_samples=10, **mog_kwargs):
        import sklearn.mixture
        if mog_kwargs.get('covariance_type', 'full') != 'diag':
            raise NotImplementedError("the GPU sampler implements covariance_type='diag' (sample_pipeline default)")
        self.mu, self.logvar = mu, logvar
        self.N, self.D = mu.shape
        self.z = torch.cat([mu + (0.5 * logvar).exp() * torch.randn_like(logvar) for _ in range(z_num_samples)], dim=0)
        self.n_components = n_components
        import cfg
        if str(getattr(cfg.b200, 'q_fit', 'sklearn')) == 'device':
            # EM on the GPU (cpg_b200.fit: sklearn's EM arithmetic in fp64; started from K data points instead of
            # sklearn's host k-means, so it reaches a different local optimum of the same likelihood)
            from cpg_b200 import fit
            self.mog = fit.gmm_fit_diag(self.z.float().to(_device()), n_components, tol=mog_kwargs.get('tol', 1e-3),
                                        max_iter=mog_kwargs.get('max_iter', 100), reg_covar=mog_kwargs.get('reg_covar', 1e-6),
                                        seed=int(cfg.seed))
        else:
            self.mog = sklearn.mixture.GaussianMixture(n_components=n_components, **mog_kwargs)
            self.mog.fit(self.z.cpu().numpy())
        self._gmm = None
        print('mog-{}. Converged: {} in {} iters, log likelihood lower bound: {:.4f}'.format(
            n_components, self.mog.converged_, self.mog.n_iter_, self.mog.lower_bound_))

    @classmethod
    def from_fitted(cls, mog):
        """Wrap an already fitted sklearn GaussianMixture (diag)."""
        self = cls.__new__(cls)
        self.mog, self.n_components, self._gmm = mog, mog.n_components, None
        self.D = mog.means_.shape[1]
        return self

    def _gmm_device(self, dev):
        if self._gmm is None or self._gmm.device != dev:
            self._gmm = sampling.GmmDevice(self.mog.weights_, self.mog.means_, self.mog.covariances_, dev)
        return self._gmm

    def logpdf(self, x):
        assert x.dim() == 1, 'expecting  single sample'
        return float(self.logpdf_batch(x.view(1, -1))[0])

    def logpdf_batch(self, x):
        dev = _device()
        return sampling.gmm_logpdf(self._gmm_device(dev), x.float().to(dev)).cpu().numpy()

    def sample(self, n_samples):
        """Host draw through sklearn / numpy's global RNG (reference density_modeling.py:79-80)."""
        if not hasattr(self.mog, 'sample'):                 # fitted on the device: same draw procedure, numpy global stream
            m = self.mog
            n_k = np.random.multinomial(n_samples, m.weights_)
            X = np.vstack([mean + np.random.standard_normal((int(k), mean.shape[0])) * np.sqrt(cov)
                           for mean, cov, k in zip(m.means_, m.covariances_, n_k)])
            return torch.from_numpy(X).float()
        return torch.from_numpy(self.mog.sample(n_samples)[0]).float()


def evaluate_nll(q, points):
    """-> (nll under Q, nll under the prior) of z = mu + exp(.5 lv) * eps with ONE scalar eps per point
    (reference density_modeling.py:99-108), batched on the GPU."""
    mu, lv = points
    n = mu.shape[0]
    eps = torch.tensor([torch.randn(1).item() for _ in range(n)], dtype=mu.dtype)
    z = mu + (0.5 * lv).exp() * eps[:, None]
    dev = _device()
    llq = q.logpdf_batch(z)
    llp = sampling.prior_logpdf(z.float().to(dev)).cpu().numpy()
    return -float(llq.sum()) / n, -float(llp.sum()) / n
