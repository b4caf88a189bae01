"""Aggregate an ncu launch list (`--metrics gpu__time_duration.sum --csv`) per kernel:
python tools/launch_summary.py gpurun_out/launches.csv [out.csv]"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki].split('(')[0].replace('void ', '')
    c = agg.setdefault(name, [0, 0.0])
    c[0] += 1
    c[1] += float(r[vi].replace(',', ''))
tot = sum(v[1] for v in agg.values())
w = csv.writer(open(sys.argv[2], 'w') if len(sys.argv) > 2 else sys.stdout)
w.writerow(['kernel', 'launches', 'total_us', 'avg_us', 'share_of_listed_time'])
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    w.writerow([k, n, '%.1f' % (t / 1e3), '%.2f' % (t / n / 1e3), '%.4f' % (t / tot)])
