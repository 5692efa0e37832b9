"""Summarise profiles/ncu_quick.sh output: per kernel (name + grid) duration, instructions, IPC-ish numbers."""
import csv, re, sys, collections
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = collections.OrderedDict()
for r in csv.DictReader(lines):
    key = int(r['ID'])
    d = rows.setdefault(key, {'name': r['Kernel Name'], 'grid': r['Grid Size']})
    v = float(r['Metric Value'].replace(',', ''))
    u = r['Metric Unit']
    if r['Metric Name'] == 'gpu__time_duration.sum':
        v = v / 1000. if u in ('nsecond', 'ns') else (v * 1000. if u in ('msecond', 'ms') else v)
    if 'bytes' in r['Metric Name']:
        v = v * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1, 'Gbyte': 1e3}.get(u, 1)
    d[r['Metric Name']] = v
agg = collections.OrderedDict()
for k, d in rows.items():
    short = re.sub(r'\(.*', '', d['name']).replace('cdra::v2::', '').replace('cdra::', '').replace('void ', '')
    key = (short[:46], d['grid'])
    a = agg.setdefault(key, [0, 0., 0., 0., 0., 0., 0.])
    a[0] += 1; a[1] += d.get('gpu__time_duration.sum', 0); a[2] += d.get('smsp__inst_executed.sum', 0)
    a[3] += d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0); a[4] += d.get('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 0)
    a[5] += d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0); a[6] = d.get('launch__registers_per_thread', 0)
tot = sum(a[1] for a in agg.values())
print(f'{"kernel":46s} {"grid":14s} {"n":>3s} {"us/launch":>9s} {"total_us":>9s} {"Minstr":>7s} {"issue%":>6s} {"Mwf_smem":>8s} {"dramMB":>7s} regs')
for (k, g), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{k:46s} {g:14s} {a[0]:3d} {a[1]/a[0]:9.1f} {a[1]:9.1f} {a[2]/a[0]/1e6:7.2f} {a[3]/a[0]:6.1f} {a[4]/a[0]/1e6:8.2f} {a[5]/a[0]:7.1f} {int(a[6])}')
print('total us', tot)
