"""CPU restatement of the peptide post-processing of a sampling round.  TEST INFRASTRUCTURE ONLY.

* drop_duplicates: pandas `Series.drop_duplicates()` as called by the reference (sample_pipeline.py:312-313) --
  keep the first occurrence of every distinct peptide.
* H / uH / charge: modlamp's GlobalAnalysis.calc_H / calc_uH / calc_charge as called by compute_modlamp
  (sample_pipeline.py:210-218).  modlamp is a third-party dependency absent from /root/reference and from this
  image (amp_gen.yml lists it without a version): its published algorithm is restated --
    PeptideDescriptor.calculate_global(window=1000, modality='max')  -> mean of the scale values (window >= len)
    PeptideDescriptor.calculate_moment(window=1000, angle=100, modality='max')
        -> sqrt((sum h_i cos(i a))^2 + (sum h_i sin(i a))^2) / len, a = 100 deg, i = 0..len-1
    GlobalDescriptor._charge(ph=7.0, amide=True) -> sum of partial charges 10^pK/(10^pK+10^pH) (N-terminus, K, R, H)
        minus 10^pH/(10^pK+10^pH) (C-terminus with pK 15 when amidated, D, E, C, Y), rounded to 3 decimals
  PARITY UNPINNED against modlamp itself (no golden vectors exist and the package cannot be imported here); the
  moment / hydrophobicity formulas are pinned against the reference's own PeptideEvaluator.calculate_moment /
  assign_hydrophobicity (evals/peptide_evals.py:73-105, restated below because the module imports Bio).
"""
import math

AA = 'ACDEFGHIKLMNPQRSTVWY'
EISENBERG = dict(A=0.62, C=0.29, D=-0.90, E=-0.74, F=1.19, G=0.48, H=-0.40, I=1.38, K=-1.50, L=1.06, M=0.64,
                 N=-0.78, P=0.12, Q=-0.85, R=-2.53, S=-0.18, T=-0.05, V=1.08, W=0.81, Y=0.26)
# evals/peptide_evals.py:18-25
EISENBERG_NORM = {'A': 0.25, 'R': -1.80, 'N': -0.64, 'D': -0.72, 'C': 0.04, 'Q': -0.69, 'E': -0.62, 'G': 0.16,
                  'H': -0.40, 'I': 0.73, 'L': 0.53, 'K': -1.10, 'M': 0.26, 'F': 0.61, 'P': -0.07, 'S': -0.26,
                  'T': -0.18, 'W': 0.37, 'Y': 0.02, 'V': 0.54}


def drop_duplicates_first(rows):
    """-> (first_index list, is_first list) for a list of hashable rows."""
    seen, first, flag = {}, [], []
    for i, r in enumerate(rows):
        key = tuple(r)
        j = seen.setdefault(key, i)
        first.append(j)
        flag.append(1 if j == i else 0)
    return first, flag


def assign_hydrophobicity(sequence, scale):            # evals/peptide_evals.py:73-88
    return [scale[aa] for aa in sequence]


def calculate_moment(array, angle=100):                 # evals/peptide_evals.py:90-105
    sum_cos, sum_sin = 0.0, 0.0
    for i, hv in enumerate(array):
        rad_inc = ((i * angle) * math.pi) / 180.0
        sum_cos += hv * math.cos(rad_inc)
        sum_sin += hv * math.sin(rad_inc)
    return math.sqrt(sum_cos ** 2 + sum_sin ** 2) / len(array)


def charge(sequence, ph=7.0, amide=True):
    pos = {'Nterm': 9.38, 'K': 10.67, 'R': 12.10, 'H': 6.04}
    neg = {'Cterm': 15.0 if amide else 2.15, 'D': 3.71, 'E': 4.15, 'C': 8.14, 'Y': 10.10}
    content = {a: sequence.count(a) for a in AA}
    content['Nterm'] = content['Cterm'] = 1
    p = sum(content[a] * 10.0 ** pk / (10.0 ** pk + 10.0 ** ph) for a, pk in pos.items())
    n = sum(content[a] * 10.0 ** ph / (10.0 ** pk + 10.0 ** ph) for a, pk in neg.items())
    return round(p - n, 3)


def descriptors(sequence, scale=EISENBERG):
    """(H, uH, charge) of one peptide string; NaN H / uH for the empty string."""
    if not sequence:
        return float('nan'), float('nan'), charge(sequence)
    h = assign_hydrophobicity(sequence, scale)
    return sum(h) / len(h), calculate_moment(h, 100), charge(sequence)
