"""Host-side helpers with the reference's names (utils.py of IBM/controlled-peptide-generation):
`anneal` / `interpolate` schedule the KL/MMD weight (utils.py:51-61), the rest are small file
utilities used by the training / sampling drivers."""
import functools
import operator
import os


def check_dir_exists(fn):
    """Create the directory part of `fn` if needed (reference utils.py:64-67)."""
    d = os.path.dirname(fn)
    if d and not os.path.isdir(d):
        os.makedirs(d, exist_ok=True)


def interpolate(start_val, end_val, start_iter, end_iter, current_iter):
    """Piecewise-linear schedule: start_val before start_iter, end_val from end_iter on."""
    if current_iter < start_iter:
        return start_val
    if current_iter >= end_iter:
        return end_val
    frac = (current_iter - start_iter) / (end_iter - start_iter)
    return start_val + (end_val - start_val) * frac


def anneal(cfgan, it):
    """Value at iteration `it` of a cfg schedule Bunch(start=Bunch(val, iter), end=Bunch(val, iter))."""
    return interpolate(cfgan.start.val, cfgan.end.val, cfgan.start.iter, cfgan.end.iter, it)


def prod(iterable):
    return functools.reduce(operator.mul, iterable, 1)


def write_gen_samples(samples, fn, c_lab=None):
    """One generated sequence per line; with labels, 'label: y' lines precede each sample."""
    check_dir_exists(fn)
    with open(fn, 'w+') as fh:
        if c_lab is None:
            print('Saving %d samples without labels' % len(samples))
            fh.write('\n'.join(samples) + '\n')
        else:
            assert c_lab.nelement() == len(samples), 'sizes dont match'
            print('Saving %d samples with labels' % len(samples))
            for y, s in zip(c_lab, samples):
                fh.write('label: {}\n{}\n'.format(y, s))


def save_vocab(vocab, fn):
    check_dir_exists(fn)
    with open(fn, 'w', encoding='utf-8') as fh:
        for word, ix in vocab.stoi.items():
            fh.write('%s %d\n' % (word, ix))
    print('Saved vocab to ' + fn)


def scale_and_clamp(dist, w, clamp_val=None):
    scaled = dist * w
    return clamp_val if (clamp_val and scaled > clamp_val) else scaled
