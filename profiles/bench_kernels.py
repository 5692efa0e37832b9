"""Print the per-kernel table of a bench.py JSON line (ms per step)."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
bk = d['roofline']['by_kernel']
n = None
tot = 0
rows = []
for k, v in bk.items():
    rows.append((v['ms'], v['launches'], k, v.get('GBps')))
    tot += v['ms']
scale = d['ms_per_step'] / tot if tot else 1
print('ms_per_step', d['ms_per_step'], 'value', d['value'], 'profiled total', tot)
for ms, n, k, g in sorted(rows, reverse=True)[:24]:
    print(f'{k[:60]:60s} n={n:4d} ms/step~{ms*scale:7.3f} share={ms/tot:.3f} GBps={g}')
