"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv ...`)
per kernel -> markdown.  usage: python tools/ncu_launches.py gpurun_out/X.csv profiles/rNN_launches.md "<command profiled>" """
import collections, csv, re, sys

src, dst, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else '')
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr, data = rows[hi], rows[hi + 1:]
ci = {h: i for i, h in enumerate(hdr)}
NOT_STEP = ('k_class', 'k_decode', 'k_compact', 'k_dedup', 'k_peptide', 'k_gmm', 'k_fill', 'k_regen', 'k_gather', 'k_score',
            'k_make_zc', 'k_sgemm_pipe', 'k_prior', 'k_feed', 'k_soft', 'k_sample')


def short(name):
    m = re.search(r'(k_\w+)', name)
    s = m.group(1) if m else name[:60]
    m2 = re.search(r'Cfg<(?:\(int\))?(\d+)', name)
    if m2 and s.startswith(('k_gru_fwd_tc', 'k_gru_bwd_fused', 'k_gru_bwd_tc')):
        s += '<dec>' if m2.group(1) == '104' else '<enc>'
    return s


agg = collections.OrderedDict()
for r in data:
    if len(r) < len(hdr):
        continue
    a = agg.setdefault(short(r[ci['Kernel Name']]), [0, 0.0])
    a[0] += 1
    a[1] += float(r[ci['Metric Value']]) / 1e3
step_k = [k for k in agg if k.startswith('k_') and not k.startswith(NOT_STEP)]
stot = sum(agg[k][1] for k in step_k)
out = ['# ncu launch list: `ncu --metrics gpu__time_duration.sum --clock-control none` over `%s`' % cmd, '',
       'Cold-cache launches serialised by the profiler: compare SHARES with the CUDA-event shares of the bench line, not absolutes.',
       'share = kernel total / total of the WAE-iteration kernels (the other rows are the CLaSS / decode legs of the same command).', '',
       '| kernel | launches | total us | mean us | share of the iteration kernels |', '|---|---:|---:|---:|---:|']
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append('| `%s` | %d | %.1f | %.1f | %s |' % (k, c, t, t / c, '%.3f' % (t / stot) if k in step_k else '-'))
open(dst, 'w').write('\n'.join(out) + '\n')
print('\n'.join(out[7:24]))
