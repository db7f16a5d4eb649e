"""CLaSS sampling driver with the reference's function names (sample_pipeline.py of
IBM/controlled-peptide-generation): rounds of  rejection_sample -> beam decode -> dedup  until enough
accepted samples exist.  The per-draw work runs on the GPU (density_modeling.py / models/model.py of
this package); this file is the thin host loop around it.

Differences from the reference, all opt-in or forced by missing third-party packages in this image:
  * `decode_accepted_only=True` decodes only the accepted z (BASELINE.json config 5); the default
    decodes every draw like the reference (sample_pipeline.py:200-201).
  * the h5 "states" files / torchtext loader / modlamp descriptors are outside the hot path
    (SURVEY.md 8f): `fit_Q` and `build_clfZ_from_encodings` take encodings as tensors, and
    `compute_modlamp` is applied only when modlamp is importable.
"""
import logging

import numpy as np
import pandas as pd
import torch

from density_modeling import mogQ

LOG = logging.getLogger('GenerationAPI')
Q_CLASS = mogQ
Q_KWARGS = {'n_components': 100, 'z_num_samples': 10, 'covariance_type': 'diag'}


def decode_from_z(z, model, dataset):
    """Beam-search decode (beam 5, hypothesis 0) in chunks of 1024 (reference :129-139)."""
    sall = []
    for zchunk in torch.split(z, 1024):
        s, _, _ = model.generate_sentences(zchunk.size(0), zchunk, sample_mode='beam', beam_size=5)
        sall += [hyps[0] for hyps in s]
    return dataset.idx2sentences(sall, print_special_tokens=False)


def get_encodings_from_dataloader(model, tokens_batches):
    """mu, logvar of every batch (reference :49-70 keeps only the encoder outputs)."""
    mus, lvs = [], []
    for tokens in tokens_batches:
        mu, lv = model.forward_encoder(tokens.to(model._param_device()))
        mus.append(mu.cpu())
        lvs.append(lv.cpu())
    return torch.cat(mus), torch.cat(lvs)


def fit_Q(mu, logvar, **overrides):
    kw = dict(Q_KWARGS)
    kw.update(overrides)
    return Q_CLASS(mu, logvar, **kw)


def build_clfZ_from_encodings(zpos_mu, zneg_mu):
    """LogisticRegression(lbfgs, 200) between attr=1 and attr=0 encodings (reference build_clfZ :169-192,
    minus the h5 query that produces the two sets)."""
    from sklearn.linear_model import LogisticRegression
    X = torch.cat([zpos_mu, zneg_mu], dim=0).numpy()
    Y = torch.cat([torch.ones(zpos_mu.shape[0]), torch.zeros(zneg_mu.shape[0])], dim=0).numpy()
    clf = LogisticRegression(solver='lbfgs', max_iter=200)
    clf.fit(X, Y)
    LOG.info('Fitted LogReg classifier in z-space: %d pos, %d neg, train accuracy=%.5f',
             zpos_mu.shape[0], zneg_mu.shape[0], clf.score(X, Y))
    return clf


def score_clfZ(clf, z):
    return clf.predict_proba(z.numpy())[:, 1]


def get_new_samples(model, dataset, Q, n_samples, decode_accepted_only=False, mode='philox'):
    """One round: rejection-sample z, decode, tabulate (reference :195-207)."""
    samples_z, scores_z, accept_z = Q.rejection_sample(n_samples=n_samples, mode=mode)
    if decode_accepted_only:
        peptides = np.full(n_samples, None, dtype=object)
        idx = np.nonzero(accept_z)[0]
        if idx.size:
            peptides[idx] = decode_from_z(samples_z[torch.from_numpy(idx)], model, dataset)
        peptides = list(peptides)
    else:
        peptides = decode_from_z(samples_z, model, dataset)
    return pd.DataFrame({'peptide': peptides, 'z': [tuple(z.tolist()) for z in samples_z], 'accept_z': accept_z,
                         **scores_z})


def compute_modlamp(df):
    try:
        from modlamp.analysis import GlobalAnalysis
    except Exception:  # noqa: BLE001  (modlamp is not installed in this image; descriptors are a "next" row)
        return df
    ana = GlobalAnalysis(df.peptide.str.replace(' ', ''))
    ana.calc_H(); ana.calc_uH(); ana.calc_charge()
    df.loc[:, 'H'], df.loc[:, 'uH'], df.loc[:, 'charge'] = ana.H[0], ana.uH[0], ana.charge[0]
    return df


def one_sampling_round(model, dataset, Q, n_samples_per_round, **kw):
    df = get_new_samples(model, dataset, Q, n_samples_per_round, **kw)
    df = compute_modlamp(df)
    df['accept'] = df['accept_z']
    return df


def run_sampling(model, dataset, Q, n_samples_per_round=5000, n_samples_acc=100, max_rounds=1000, **kw):
    """The while-loop of the reference's main() (:303-322): sample until n_samples_acc accepted, dropping
    duplicate peptides within and across rounds."""
    samples = pd.DataFrame(columns=['peptide'])
    rounds = 0
    while (len(samples) < n_samples_acc or samples['accept'].sum() < n_samples_acc) and rounds < max_rounds:
        rounds += 1
        new = one_sampling_round(model, dataset, Q, n_samples_per_round, **kw)
        new = new[new.peptide.notna()]
        new = new.loc[new.peptide.drop_duplicates().index]
        new = new[~new['peptide'].isin(samples['peptide'])]
        samples = pd.concat([samples, new], ignore_index=True, sort=False)
        LOG.info('round %d: %d rows, %d accepted', rounds, len(samples), int(samples['accept'].sum()))
    return samples
