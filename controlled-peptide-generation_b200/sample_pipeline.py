"""CLaSS sampling pipeline with the reference's entry points (sample_pipeline.py of
IBM/controlled-peptide-generation): fit Q_xi(z) on the encodings, fit the z-space attribute classifiers,
then rounds of  rejection_sample -> beam decode -> dedup  until enough accepted samples exist.
Same function names, arguments and CLI flags: get_encodings(_from_dataloader/_from_states), fitQ_and_test,
decode_from_z, save_samples, score_clfZ, build_clfZ, get_new_samples, compute_modlamp, one_sampling_round,
main(args), `python sample_pipeline.py --Q_n_components ... --n_samples_acc ...` (reference :327-361).

The per-draw work runs on the GPU (density_modeling.py / models/model.py of this package):
  * draws, classifier scores, accept test          cpg_class_sample (Philox)   [density_modeling.mogQ]
  * compaction of the accepted z + beam decode      cpg_compact_rows + cpg_beam_decode, no host round trip
  * peptide dedup + H / uH / charge descriptors     cpg_dedup_tokens + cpg_peptide_descriptors (cpg_b200.peptides)

Differences from the reference, all opt-in or forced by packages this image lacks:
  * cfg.b200.decode_accepted_only = True decodes only the accepted z (BASELINE.json config 5); the default decodes
    every draw like the reference (:200-201).
  * states files are `.h5` when h5py is importable, else `.npz` with the same datasets (cpg_b200/states.py).
  * modlamp is replaced by the package's own Eisenberg-scale kernels (same formulas); the torchtext
    AttributeDataLoader and the PeptideEvaluator are only imported by main() when no dataset is passed in.
"""
import argparse
import datetime
import json
import logging
import os
import pprint
import sys
from collections import OrderedDict

import numpy as np
import pandas as pd
import torch

import cfg
from cpg_b200 import states as states_io
from density_modeling import evaluate_nll, mogQ

LOG = logging.getLogger('GenerationAPI')
pp = pprint.PrettyPrinter(indent=2, depth=1)

Q_CLASS = mogQ
Q_KWARGS = {'n_components': None, 'z_num_samples': 10, 'covariance_type': None}


def get_encodings(query, split, model=None, dataloader=None):
    if model and dataloader:
        return get_encodings_from_dataloader(query, split, model, dataloader)
    return get_encodings_from_states(query, split)


def get_encodings_from_dataloader(query, split, model, dataloader):
    """mu, logvar of the amp-positive sequences of `split` (reference :49-70).  `dataloader` is the
    reference's AttributeDataLoader (its torchtext subset iterators are used when present) or any object with
    `encoding_batches(split, query)` yielding int64 token batches."""
    assert query == {'amp': 1}, 'only support this right now (reference sample_pipeline.py:50-53)'
    if hasattr(dataloader, 'encoding_batches'):
        batches = dataloader.encoding_batches(split, query)
    else:
        specs = {'get_encoding': {'subset': ['split=' + split, 'amp=amp_posc,amp_posnc'], 'repeat': False}}
        iterators, _ = dataloader.dataset.get_subset_iterators(specs, cfg.vae.batch_size, torch.device('cpu'))
        batches = (b.text for b in iter(iterators['get_encoding']))
    mus, logvars = [], []
    dev = model._param_device()
    for tokens in batches:
        with torch.no_grad():
            mu, logvar = model.forward_encoder(tokens.to(dev))          # what model(..., sample_z='max') keeps
        mus.append(mu.detach().cpu())
        logvars.append(logvar.detach().cpu())
    return torch.cat(mus, dim=0), torch.cat(logvars, dim=0)


def get_encodings_from_states(query, split):
    """float16 mu / logvar rows of `states_{split}_{n_iter}` whose labels match `query` (reference :73-92)."""
    attr_to_colix = {k: i for i, (k, _) in enumerate(cfg.attributes)}
    st = states_io.read_states(states_io.states_basename(cfg.savepath, split, cfg.vae.n_iter))
    mu, logvar = torch.from_numpy(st['mu'][:]).double(), torch.from_numpy(st['logvar'][:]).double()
    lab = torch.from_numpy(st['label'][:])
    sel = torch.ones(lab.shape[0], dtype=torch.bool)
    for attr_name, val in query.items():
        sel &= lab[:, attr_to_colix[attr_name]] == val
    return mu[sel], logvar[sel]


def fitQ_and_test(QClass, QKwargs, Q_select={}, negative_select={}, model=None, dataloader=None):
    """Fit Q_xi^a(z) on the encodings selected by `Q_select`; NLL under Q and under the prior on the train /
    held-out selections (reference :95-126)."""
    if model and dataloader:
        mu, logvar = get_encodings_from_dataloader(query=Q_select, split='train,val', model=model, dataloader=dataloader)
    else:
        mu, logvar = get_encodings_from_states(query=Q_select, split='train')
    Q_xi_a = QClass(mu, logvar, **QKwargs)
    LOG.info('Fitted {}  {} on selection {}'.format(QClass.__name__, str(QKwargs), str(Q_select)))
    metrics = OrderedDict()
    for name, points in (('a,tr', get_encodings_from_states(split='train', query=Q_select)),
                         ('a,hld', get_encodings_from_states(split='test', query=Q_select))):
        metrics[name] = evaluate_nll(Q_xi_a, points)
    return Q_xi_a, metrics


def decode_from_z(z, model, dataset):
    """Beam-search decode (beam 5, hypothesis 0) in chunks of 1024 (reference :129-139)."""
    sall = []
    LOG.info('Decoder decoding: beam search')
    for zchunk in torch.split(z, 1024):
        s, _, _ = model.generate_sentences(zchunk.size(0), zchunk, sample_mode='beam', beam_size=5)
        sall += [hypotheses[0] for hypotheses in s]
    return dataset.idx2sentences(sall, print_special_tokens=False)


def save_csv_pkl(samples, fn):
    samples.drop(columns='z').to_csv(fn + '.csv', index_label='idx')
    samples.to_pickle(fn + '.pkl')


def save_samples(samples, basedir, fn_prefix):
    outfn = os.path.join(basedir, fn_prefix) + '_{}'.format(datetime.datetime.now().isoformat().split('T')[0])
    with open(outfn + '.plain.txt', 'w') as fh:
        fh.write(samples['peptide'].to_string(index=False))
    save_csv_pkl(samples, outfn)
    LOG.info('Full sample list written to {}.pkl/csv'.format(outfn))
    accepted = samples[samples.accept.astype(bool)]
    accepted_fn = '{}.accepted.{}'.format(outfn, len(accepted))
    save_csv_pkl(accepted, accepted_fn)
    LOG.info('Accepted sample list written to {}.pkl/csv'.format(accepted_fn))
    return outfn


def score_clfZ(clf, z):
    return clf.predict_proba(z.numpy())[:, 1]


def fit_clfZ(zpos_mu, zneg_mu):
    """LogisticRegression(lbfgs, 200) between attr=1 and attr=0 encodings (the fit inside build_clfZ)."""
    X = torch.cat([zpos_mu, zneg_mu], dim=0).numpy()
    Y = torch.cat([torch.ones(zpos_mu.shape[0]), torch.zeros(zneg_mu.shape[0])], dim=0).numpy()
    if str(getattr(cfg.b200, 'clf_fit', 'sklearn')) == 'device':
        from cpg_b200.fit import DeviceLogisticRegression       # same objective, Newton steps with GPU statistics
        clf = DeviceLogisticRegression()
    else:
        from sklearn.linear_model import LogisticRegression
        clf = LogisticRegression(solver='lbfgs', max_iter=200)
    clf.fit(X, Y)
    LOG.info('num samples: {} pos, {} neg. train accuracy={:.5f}'.format(zpos_mu.shape[0], zneg_mu.shape[0], clf.score(X, Y)))
    return clf


def build_clfZ(attr):
    """sklearn logistic regression between attr=1 and attr=0 train encodings (labels -1 / 0 / 1 = na / neg / pos),
    reference :169-192."""
    zpos_mu, _ = get_encodings_from_states(query={attr: 1}, split='train')
    zneg_mu, _ = get_encodings_from_states(query={attr: 0}, split='train')
    clf = fit_clfZ(zpos_mu, zneg_mu)
    LOG.info('Fitted LogReg classifier in z-space, on attr={}.'.format(attr))
    return clf


def get_new_samples(model, dataset, Q, n_samples, decode_accepted_only=None, mode='philox'):
    """One round: rejection-sample z, decode, tabulate (reference :195-207)."""
    if decode_accepted_only is None:
        decode_accepted_only = bool(getattr(cfg.b200, 'decode_accepted_only', False))
    if decode_accepted_only and mode == 'philox' and hasattr(Q, 'rejection_sample_decode'):
        return Q.rejection_sample_decode(n_samples, model, dataset)
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
    """H (Eisenberg hydrophobicity), uH (hydrophobic moment, window 11, 100 deg), charge (pH 7, amide False) --
    the three modlamp GlobalAnalysis descriptors of the reference (:210-218), computed by the package's kernels."""
    from cpg_b200 import peptides
    seqs = df.peptide.fillna('').str.replace(' ', '')
    H, uH, charge = peptides.descriptors_from_strings(list(seqs))
    df.loc[:, 'H'], df.loc[:, 'uH'], df.loc[:, 'charge'] = H, uH, charge
    return df


def one_sampling_round(model, dataset, Q, n_samples_per_round, **kw):
    samples_df = get_new_samples(model, dataset, Q, n_samples_per_round, **kw)
    if 'H' not in samples_df.columns:                  # the device pipeline already attached H / uH / charge
        samples_df = compute_modlamp(samples_df)
    samples_df['accept'] = samples_df['accept_z']
    return samples_df


def get_sample_source_str():
    return ' '.join(sys.argv[1:])


def _arg(args, name, default):
    return getattr(args, name, default) if not isinstance(args, dict) else args.get(name, default)


def run_sampling(model, dataset, Q, n_samples_per_round=5000, n_samples_acc=100, max_rounds=1000, **kw):
    """The sampling loop of main() (reference :296-322): rounds until n_samples_acc accepted, dropping duplicate
    peptides within and across rounds."""
    samples = pd.DataFrame(columns=['peptide'])
    round_ix = 0

    def is_finished(df, min_accepted):
        return not (len(df) < min_accepted or df['accept'].sum() < min_accepted)

    while not is_finished(samples, n_samples_acc) and round_ix < max_rounds:
        round_ix += 1
        LOG.info('Round #{}'.format(round_ix))
        new_samples = one_sampling_round(model, dataset, Q, n_samples_per_round, **kw)
        new_samples = new_samples[new_samples.peptide.notna()]
        new_samples = new_samples.loc[new_samples.peptide.drop_duplicates().index]
        new_samples = new_samples[~new_samples['peptide'].isin(samples['peptide'])]
        samples = pd.concat([samples, new_samples], ignore_index=True, sort=False)
        dropped_num = n_samples_per_round - new_samples.shape[0]
        if dropped_num > 0:
            LOG.info('Dropped {} duplicate samples'.format(dropped_num))
        LOG.info('Q_xi(z|a) rejection sampling acceptance rate: {}/{} = {:.4f}'.format(
            samples['accept_z'].sum(), len(samples), 100.0 * samples['accept_z'].sum() / max(len(samples), 1)))
    return samples


def main(args={}, model=None, dataset=None):
    """Reference :236-324.  `model` / `dataset` may be passed in (tests, notebooks); otherwise they are loaded
    from cfg.savepath exactly as the reference does (api.load_trained_model + AttributeDataLoader)."""
    if model is None:
        from api import Vocab, get_model_and_vocab_path, get_result_for_model, load_trained_model
        MODEL_PATH, VOCAB_PATH, _ = get_model_and_vocab_path()
        LOG.info('Load model, vocab, dataloader.')
        vocab = Vocab(VOCAB_PATH)
        model = load_trained_model(MODEL_PATH, vocab.size())
        LOG.info('Loaded model succesfully.')
        metrics = get_result_for_model(MODEL_PATH, print_results=False)
        LOG.info('Model metrics:')
        pp.pprint(metrics)
    torch.manual_seed(cfg.seed)
    np.random.seed(cfg.seed)
    if dataset is None:
        from data_processing.dataset import AttributeDataLoader     # torchtext loader: not part of this package
        dataset = AttributeDataLoader(mbsize=cfg.vae.batch_size, max_seq_len=cfg.max_seq_len,
                                      device=torch.device('cpu'), attributes=cfg.attributes, **cfg.data_kwargs)

    LOG.info('Fit attribute-conditioned marginal posterior Q_xi^a(z)')
    q_kwargs = dict(Q_KWARGS)
    for k in q_kwargs:
        v = _arg(args, 'Q_' + k, None)
        if v is not None:
            q_kwargs[k] = v
    if _arg(args, 'Q_select_amppos', 0):
        Q_SELECT_QUERY, Q_NEGATIVE_QUERY = {'amp': 1}, {'amp': 0}
    else:
        Q_SELECT_QUERY, Q_NEGATIVE_QUERY = {}, {}
    full = _arg(args, 'Q_from_full_dataloader', False)
    Q, Q_xi_metrics = fitQ_and_test(Q_CLASS, q_kwargs, Q_SELECT_QUERY, Q_NEGATIVE_QUERY, model if full else None,
                                    dataset if full else None)
    LOG.info('Q Fit metrics: ')
    print(json.dumps(Q_xi_metrics, indent=4))

    z_clfs = OrderedDict()
    for attr in ['amp', 'tox']:
        z_clfs[attr] = build_clfZ(attr)
    Q.init_attr_classifiers(z_clfs, clf_targets={'amp': 1, 'tox': 0})

    # ---- setup done, sampling below
    samples = run_sampling(model, dataset, Q, n_samples_per_round=_arg(args, 'n_samples_per_round', 5000),
                           n_samples_acc=_arg(args, 'n_samples_acc', 100))
    LOG.info('     - full filter pipeline accepted: {}/{} = {:.4f}'.format(
        samples['accept'].sum(), len(samples), 100.0 * samples['accept'].sum() / max(len(samples), 1)))
    save_samples(samples, cfg.savepath, _arg(args, 'samples_outfn_prefix', 'samples'))
    return samples


def build_parser():
    parser = argparse.ArgumentParser(argument_default=argparse.SUPPRESS, description='Override config float & string values')
    cfg._cfg_import_export(parser, cfg, mode='fill_parser')
    parser.add_argument('--QClass', default='mogQ')
    parser.add_argument('--Q_n_components', type=int, default=100, help='mog num components for Q model')
    parser.add_argument('--Q_covariance_type', default='diag', help='mog Q covariance type (the GPU sampler builds diag)')
    parser.add_argument('--n_samples_per_round', type=int, default=5000, help='number of samples to generate & evaluate.')
    parser.add_argument('--n_samples_acc', type=int, default=100, help='number of accepted samples to collect.')
    parser.add_argument('--samples_outfn_prefix', default='samples', help='prefix of the .txt .csv .pkl outputs')
    parser.add_argument('--Q_select_amppos', type=int, default=0, help='select amp positive to fit Q_xi or not.')
    parser.add_argument('--Q_from_full_dataloader', action='store_true', default=False,
                        help='to fit Q_z, select from full dataloader')
    return parser


if __name__ == '__main__':
    logging.basicConfig(format='%(asctime)s %(message)s', datefmt='%m/%d/%Y %I:%M:%S %p', level=logging.INFO)
    LOG.info('Sample pipeline. Fit Q_xi(z), Sample from it, score samples.')
    cli_args = build_parser().parse_args()
    cfg._override_config(cli_args, cfg)
    cfg._update_cfg()
    cfg._print(cfg)
    main(cli_args)
