"""The "states" files of the reference pipeline: per split, the encoder outputs of up to max_examples
sequences with their labels (vis/scripts/build_index.py:30-118 writes them, sample_pipeline.py:73-92 reads
them).  Same datasets, same dtypes -- src int [N, L], z / mu / logvar FLOAT16 [N, 100], label int [N, n_attr],
split int [N, 1] -- as `states_{split}_{n_iter}.h5` (gzip-9, chunks (10, 5) like the reference) when h5py is
importable, else `states_{split}_{n_iter}.npz` with the same keys (this image has no h5py).  The float16
rounding matters: the reference fits Q(z) and the z-space classifiers on the ROUNDED mu / logvar.

Extraction runs the encoder kernels (cpg_wae_encode via RNN_VAE.forward_encoder); z = mu as with
sample_z='max' (build_index.py:98-99).
"""
import os

import numpy as np
import torch

SPLIT_ENCODING = {'train': 0, 'val': 1, 'test': 2}
KEYS = ('src', 'z', 'mu', 'logvar', 'label', 'split')


def states_basename(base_folder, split, n_iter):
    return os.path.join(base_folder, 'states_{}_{}'.format(split, n_iter))


def _have_h5py():
    try:
        import h5py  # noqa: F401
        return True
    except Exception:  # noqa: BLE001
        return False


def write_states(path_base, src, z, mu, logvar, label, split):
    """Arrays / tensors -> path_base + ('.h5' | '.npz').  Returns the file written."""
    def arr(x, dt):
        return np.ascontiguousarray(torch.as_tensor(x).detach().cpu().numpy().astype(dt))
    data = {'src': arr(src, np.int64), 'z': arr(z, np.float16), 'mu': arr(mu, np.float16),
            'logvar': arr(logvar, np.float16), 'label': arr(label, np.int64),
            'split': arr(split, np.int64).reshape(-1, 1)}
    n = data['src'].shape[0]
    assert all(v.shape[0] == n for v in data.values()), 'row counts differ'
    for ext in ('.h5', '.npz'):
        if os.path.isfile(path_base + ext):
            os.remove(path_base + ext)
    if _have_h5py():
        import h5py
        with h5py.File(path_base + '.h5', 'w') as f:
            for k, v in data.items():
                chunks = (min(10, max(n, 1)), min(1 if k == 'split' else 5, v.shape[1]))
                f.create_dataset(k, data=v, maxshape=(None, None), chunks=chunks, compression='gzip',
                                 compression_opts=9)
        return path_base + '.h5'
    np.savez_compressed(path_base + '.npz', **data)
    return path_base + '.npz'


def read_states(path_base):
    """-> dict of numpy arrays (KEYS) from the .h5 (h5py) or .npz file."""
    if os.path.isfile(path_base + '.h5'):
        import h5py
        with h5py.File(path_base + '.h5', 'r') as f:
            return {k: f[k][:] for k in KEYS if k in f}
    if os.path.isfile(path_base + '.npz'):
        with np.load(path_base + '.npz') as f:
            return {k: f[k] for k in f.files}
    raise FileNotFoundError(path_base + '.{h5,npz}: need dumped states, run extract_states / static_eval first')


def extract_states(model, batches, base_folder, split, n_iter, max_examples=20000):
    """batches: iterable of (tokens int64 [b, L], labels int [b, n_attr]).  Encodes on the GPU, keeps at most
    max_examples rows, writes the states file of `split`.  Returns its path."""
    src, mus, lvs, labs = [], [], [], []
    n = 0
    dev = model._param_device()
    for tokens, labels in batches:
        with torch.no_grad():
            mu, lv = model.forward_encoder(tokens.to(dev))
        src.append(tokens.cpu())
        mus.append(mu.to(torch.float16).cpu())
        lvs.append(lv.to(torch.float16).cpu())
        labs.append(torch.as_tensor(labels).cpu())
        n += tokens.shape[0]
        if n >= max_examples:
            break
    cat = lambda xs: torch.cat(xs)[:max_examples]
    src, mu, lv, lab = cat(src), cat(mus), cat(lvs), cat(labs)
    split_col = torch.full((src.shape[0], 1), SPLIT_ENCODING[split], dtype=torch.int64)
    os.makedirs(base_folder, exist_ok=True)
    return write_states(states_basename(base_folder, split, n_iter), src, mu, mu, lv, lab, split_col)
