"""Summarise an ncu report (read here, no GPU): per-kernel duration, DRAM traffic and pipe utilisation ->
profiles/<name>.md, and the per-launch DRAM traffic table bench.py reads (profiles/ncu_traffic.json).
usage: python tools/ncu_summary.py gpurun_out/X.ncu-rep profiles/rNN_name"""
import csv, io, json, os, re, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
M = [('gpu__time_duration.sum', 'us'), ('dram__bytes_read.sum', 'rd MB'), ('dram__bytes_write.sum', 'wr MB'),
     ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram %'),
     ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor %'),
     ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue %'),
     ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'xu %'),
     ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1 %'),
     ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2 %'),
     ('launch__registers_per_thread', 'regs'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block')]


def short(name):
    m = re.search(r'(k_\w+)(<[^(]*>)?', name)
    s = m.group(0) if m else name
    s = re.sub(r'\(int\)|\(bool\)|cpg::|<unnamed>::|unnamed>::', '', s)
    return s[:70]


def to_mb(v, unit):
    v = float(v)
    return v * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(unit, 1.0)


lines = ['| kernel | ' + ' | '.join(n for _, n in M) + ' |', '|---|' + '---|' * len(M)]
agg = {}
for r in data:
    name = short(r[col['Kernel Name']])
    vals = []
    for m, n in M:
        v, u = r[col[m]], units[col[m]]
        if 'bytes' in m:
            v = '%.2f' % to_mb(v, u)
        elif m == 'gpu__time_duration.sum':
            v = '%.1f' % (float(v) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(u, 1.0))
        else:
            v = ('%.1f' % float(v)) if '.' in v else v
        vals.append(v)
    lines.append('| `%s` | ' % name + ' | '.join(vals) + ' |')
    key = re.match(r'k_\w+', name).group(0) if name.startswith('k_') else name
    for stem, fmt in (('k_gru_fwd_tc', 'k_gru_fwd_%s_tc'), ('k_gru_bwd_tc', 'k_gru_bwd_%s_tc'), ('k_gru_bwd_fused', 'k_gru_bwd_%s_fused')):
        if key.startswith(stem) and 'Cfg' in name:
            key = fmt % ('dec' if '104' in name.split('Cfg')[1][:6] else 'enc')
            break
    if key == 'k_latent_fwd_tc':
        key = 'k_latent_fwd_tc<M>'
    if 'k_wgrad_tc' in key:
        key += '_dec' if '<104>' in name else '_enc'
    agg.setdefault(key, []).append(to_mb(r[col['dram__bytes_read.sum']], units[col['dram__bytes_read.sum']]) +
                                   to_mb(r[col['dram__bytes_write.sum']], units[col['dram__bytes_write.sum']]))
with open(out + '.md', 'w') as f:
    f.write('# ncu --set full --clock-control none: %s\n\n(cold-cache, serialised replays: compare shares, not absolutes)\n\n' % os.path.basename(rep))
    f.write('\n'.join(lines) + '\n')
traffic = {k: int(sum(v) / len(v) * 1e6) for k, v in agg.items()}
tpath = os.path.join(os.path.dirname(out), 'ncu_traffic.json')
old = {}
if os.path.exists(tpath):
    old = json.load(open(tpath))
old.update(traffic)
json.dump(old, open(tpath, 'w'), indent=1, sort_keys=True)
print('wrote', out + '.md', tpath)
