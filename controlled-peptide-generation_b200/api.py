"""Model / vocabulary loading helpers with the reference's names (api.py of
IBM/controlled-peptide-generation: Vocab :27-72, load_trained_model :75-96, encode_sequence :99-113,
sample_from_model :116-148, get_model_and_vocab_path :289-304, get_result_for_model :307-334), as used by
sample_pipeline.main.  The model they return runs on the B200 kernels, so it lives on a CUDA device
(the reference pins its copy to the CPU, api.py:94).
"""
import codecs
import json
import logging
import os

import torch

import cfg
from models.model import RNN_VAE

LOG = logging.getLogger('GenerationAPI')
SPECIAL_TOKENS = ('<unk>', '<pad>', '<start>', '<eos>')


class Vocab:
    """word <-> index tables read from a `word index` per line file."""

    def __init__(self, VOCAB_PATH):
        self.fix_length = cfg.max_seq_len
        self.ix2word, self.word2ix = {}, {}
        with codecs.open(VOCAB_PATH, 'r', 'utf-8') as fh:
            for line in fh:
                parts = line.split()
                if not parts:
                    continue
                word, ix = ' '.join(parts[:-1]), int(parts[-1])
                self.ix2word[ix] = word
                self.word2ix[word] = ix
        self.special_tokens = set(SPECIAL_TOKENS)
        self.special_tokens_ix = {self.word2ix[w] for w in self.special_tokens if w in self.word2ix}

    def to_ix(self, seq, fix_length=True):
        if isinstance(seq, str):
            seq = seq.split()
        elif not isinstance(seq, list):
            raise ValueError('Only strings or lists of strings accepted.')
        seq = list(seq)
        if seq[0] != '<start>':
            seq.insert(0, '<start>')
        if seq[-1] != '<eos>':
            seq.append('<eos>')
        if fix_length:
            seq += ['<pad>'] * (self.fix_length - len(seq))
        return torch.LongTensor([self.word2ix[t] for t in seq]).view(1, -1)

    def to_word(self, seq, print_special_tokens=True):
        ids = [int(s) for s in seq]
        if not print_special_tokens:
            ids = [i for i in ids if i not in self.special_tokens_ix]
        return [self.ix2word[i] for i in ids]

    def size(self):
        return len(self.ix2word)


def load_trained_model(MODEL_PATH, n_vocab, device='cuda'):
    model = RNN_VAE(n_vocab, max_seq_len=cfg.max_seq_len, **cfg.model)
    model.load_state_dict(torch.load(MODEL_PATH, map_location='cpu'), strict=False)
    model = model.to(device)
    model.eval()
    return model


def encode_sequence(model, vocab, sequence, sample_q='max'):
    mu, logvar = model.forward_encoder(vocab.to_ix(sequence).to(model._param_device()))
    if sample_q == 'max':
        return mu
    return torch.cat([model.sample_z(mu, logvar) for _ in range(sample_q)], dim=0)


def sample_from_model(model, vocab, z=None, c=None, n_samples=2, print_special_tokens=True, **sample_kwargs):
    samples, z, c = model.generate_sentences(n_samples, z=z, c=c, **sample_kwargs)
    if sample_kwargs.get('sample_mode') == 'beam':
        predictions = [[vocab.to_word(h, print_special_tokens) for h in hyps] for hyps in samples]
    else:
        predictions = [[vocab.to_word(s, print_special_tokens)] for s in samples]
    return {'predictions': predictions, 'z': z, 'c': c}


def get_model_and_vocab_path():
    """Final VAE checkpoint of the run directory (or the latest one present) and its vocabulary file."""
    base = cfg.savepath
    want = 'model_{}.pt'.format(cfg.vae.n_iter)
    files = os.listdir(base)
    if want not in files:
        its = [int(n.split('_')[1].split('.')[0]) for n in files if n.startswith('model_') and n.endswith('.pt')]
        if not its:
            raise FileNotFoundError('no model_*.pt checkpoint in ' + base)
        LOG.info('Selected model folder does not have fully trained model! Using iteration %d instead', max(its))
        want = 'model_{}.pt'.format(max(its))
    return os.path.join(base, want), os.path.join(base, 'vocab.dict'), base


def get_result_for_model(model_path, print_results=False):
    """Entry of result.json (list of per-iteration dicts) that belongs to this checkpoint, {} if none."""
    fn = os.path.join(os.path.dirname(model_path), 'result.json')
    if not os.path.isfile(fn):
        LOG.info('No results for %s found.', model_path)
        return {}
    with open(fn) as fh:
        data = json.load(fh)
    it = os.path.basename(model_path).split('.')[0].split('_')[1]
    stats = {}
    for res in data:
        if str(res.get('it')) == str(it):
            stats = res
    if print_results:
        print('Results for model {}'.format(model_path))
        print(json.dumps(stats, indent=2))
    return stats


def main(args={}):
    MODEL_PATH, VOCAB_PATH, _ = get_model_and_vocab_path()
    vocab = Vocab(VOCAB_PATH)
    load_trained_model(MODEL_PATH, vocab.size())
    LOG.info('loaded successfully.')
