"""Summarise an `ncu --page raw --csv` dump: python tools/ncu_summary.py raw.csv [out.csv]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
base = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
        'smsp__cycles_active.avg']
stalls = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
ki = hdr.index('Kernel Name')
out = []
for r in data:
    name = r[ki].split('(')[0].replace('void ', '')
    rec = {'kernel': name}
    for b in base:
        if b in hdr:
            rec[b] = r[hdr.index(b)] + ' ' + units[hdr.index(b)]
    st = sorted(((float(r[hdr.index(s)]), s[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')])
                 for s in stalls if r[hdr.index(s)] not in ('', 'n/a')), reverse=True)[:6]
    rec['top_stalls(warps per issue)'] = '; '.join('%s %.2f' % (n, v) for v, n in st)
    out.append(rec)
keys = ['kernel'] + [b for b in base if b in hdr] + ['top_stalls(warps per issue)']
w = csv.writer(open(sys.argv[2], 'w') if len(sys.argv) > 2 else sys.stdout)
w.writerow(keys)
for rec in out:
    w.writerow([rec.get(k, '') for k in keys])
