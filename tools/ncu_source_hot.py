"""Aggregate ncu per-line samples: python tools/ncu_source_hot.py report.ncu-rep [topN]"""
import csv, subprocess, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; agg = collections.OrderedDict()
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) == 2: continue
    if r[0] == 'Line No': hdr = r; continue
    if r[2] != '-': continue  # keep only the per-source-line summary rows (Address == '-')
    d = dict(zip(hdr[4:], r[4:]))
    key = (cur, int(r[0]), r[1].strip())
    s = float(d['# Samples'] or 0); ins = float(d['Instructions Executed'] or 0); thr = float(d['Thread Instructions Executed'] or 0)
    a = agg.setdefault(key, [0, 0, 0]); a[0] += s; a[1] += ins; a[2] += thr
ts = sum(a[0] for a in agg.values()); ti = sum(a[1] for a in agg.values()); tt = sum(a[2] for a in agg.values())
print('samples %d  warp-inst %.3g  avg threads/inst %.1f' % (ts, ti, tt / max(ti, 1)))
for (f, ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print('%5.1f%% smp %5.1f%% inst thr %4.1f | %-18s:%4d | %s' % (100 * a[0] / ts, 100 * a[1] / ti, a[2] / max(a[1], 1), f, ln, src[:100]))
