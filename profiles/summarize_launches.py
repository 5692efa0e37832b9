"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of the step)."""
import csv, re, sys, collections
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]; ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r'<.*', '', r[ki].split('(')[0]).replace('cdra::', '').strip()
    v = float(r[vi].replace(',', '')); v = v / 1e3 if r[ui] in ('ns', 'nsecond') else v
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f'{"kernel":44s} {"launches":>8s} {"total_us":>12s} {"share":>7s} {"avg_us":>9s}')
for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{k:44s} {c:8d} {us:12.1f} {us/tot:7.3f} {us/c:9.1f}')
print(f'{"TOTAL":44s} {sum(a[0] for a in agg.values()):8d} {tot:12.1f}')
