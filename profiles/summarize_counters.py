"""Per-kernel summary of an `ncu --metrics ... --csv` counter list: launches, average duration, DRAM GB/s, DRAM bytes per
launch, tensor-pipe activity, issue utilisation.   python profiles/summarize_counters.py <csv> [<json out>]"""
import collections, csv, json, re, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
L = collections.OrderedDict()
for r in rows[1:]:
    d = L.setdefault(r[ix['ID']], {'name': re.sub(r'\(.*', '', r[ix['Kernel Name']]).replace('void ', '').replace('cdra::v2::', '')})
    d[r[ix['Metric Name']]] = float(r[ix['Metric Value']].replace(',', ''))
agg = collections.OrderedDict()
for d in L.values():
    a = agg.setdefault(d['name'], collections.defaultdict(float)); a['n'] += 1
    a['us'] += d['gpu__time_duration.sum'] / 1e3
    a['bytes'] += d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)
    a['tensor'] += d.get('sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active', 0)
    a['dram_pct'] += d.get('dram__throughput.avg.pct_of_peak_sustained_elapsed', 0)
    a['issue'] += d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0)
out = {}
print(f'{"kernel":28s} {"n":>4s} {"avg us":>8s} {"DRAM MB/launch":>15s} {"DRAM GB/s":>10s} {"dram %":>7s} {"tensor %":>9s} {"issue %":>8s}')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['us']):
    n = a['n']
    print(f'{k:28s} {int(n):4d} {a["us"] / n:8.1f} {a["bytes"] / n / 1e6:15.1f} {a["bytes"] / a["us"] / 1e3:10.1f} {a["dram_pct"] / n:7.1f} {a["tensor"] / n:9.2f} {a["issue"] / n:8.1f}')
    out[k] = {'launches': int(n), 'avg_us': a['us'] / n, 'dram_bytes_per_launch': a['bytes'] / n, 'dram_GBps': a['bytes'] / a['us'] / 1e3,
              'dram_pct_of_peak': a['dram_pct'] / n, 'tensor_pipe_pct': a['tensor'] / n, 'issue_pct': a['issue'] / n}
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], 'w'), indent=1)
