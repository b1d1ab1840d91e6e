"""Per-source-line stall attribution from an ncu report: python tools/ncu_stall_lines.py report.ncu-rep [stall_col] [topN]
Rows are the per-line summaries of the source page; stall_col e.g. stall_long_sb, stall_wait, stall_no_inst."""
import csv, subprocess, sys, collections
path = sys.argv[1]; col = sys.argv[2] if len(sys.argv) > 2 else 'stall_long_sb'; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
cur = None; hdr = None; agg = collections.OrderedDict(); tot = collections.Counter()
for r in csv.reader(out.splitlines()):
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) == 2: continue
    if r[0] == 'Line No': hdr = r; continue
    if r[2] != '-': continue
    d = dict(zip(hdr[4:], r[4:]))
    def f(k):
        try: return float(d.get(k) or 0)
        except ValueError: return 0.0
    key = (cur, int(r[0]), r[1].strip())
    a = agg.setdefault(key, collections.Counter())
    a['smp'] += f('# Samples'); a['col'] += f(col); a['inst'] += f('Instructions Executed')
    for k in d:
        if k.startswith('stall_') and 'Not Issued' not in k: tot[k] += f(k)
ts = sum(a['smp'] for a in agg.values()); tc = sum(a['col'] for a in agg.values())
print('total samples %d; %s %d (%.1f%%)' % (ts, col, tc, 100 * tc / max(ts, 1)))
print('all stall columns:', ', '.join('%s %.1f%%' % (k[6:], 100 * v / max(ts, 1)) for k, v in tot.most_common(10)))
for (fn, ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1]['col'])[:top]:
    print('%5.1f%% of %s | %5.1f%% smp | %-16s:%4d | %s' % (100 * a['col'] / max(tc, 1), col[6:], 100 * a['smp'] / ts, fn, ln, src[:110]))
