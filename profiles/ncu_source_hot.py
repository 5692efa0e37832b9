"""Hot spots of an `ncu --page source --csv` export (SASS view): per kernel, the instructions with the most stall samples
and the shared-memory accesses with the most excess wavefronts."""
import csv, sys, re, collections
txt = open(sys.argv[1]).read()
top = int(sys.argv[2]) if len(sys.argv) > 2 else 14
blocks = re.split(r'(?m)^(?="Kernel Name")', txt)
for b in blocks:
    if not b.strip(): continue
    lines = b.strip().split('\n')
    name = next(csv.reader([lines[0]]))[1]
    rows = list(csv.DictReader(lines[1:]))
    tot = sum(int(r['# Samples']) for r in rows)
    print('=' * 20, name[:90], 'samples', tot, 'instr', len(rows))
    reasons = [k for k in rows[0] if k.startswith('stall_') and 'Not Issued' not in k]
    agg = {k: sum(int(r[k]) for r in rows) for k in reasons}
    print('  stalls:', ', '.join(f'{k[6:]} {v*100//max(tot,1)}%' for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    ops = collections.Counter()
    for r in rows: ops[r['Source'].split()[0] if r['Source'].split()[0][0] != '@' else r['Source'].split()[1]] += int(r['# Samples'])
    print('  by opcode:', ', '.join(f'{k} {v*100//max(tot,1)}%' for k, v in ops.most_common(12)))
    for i, r in sorted(enumerate(rows), key=lambda ir: -int(ir[1]['# Samples']))[:top]:
        why = max(reasons, key=lambda k: int(r[k]))
        print(f'   #{i:5d} {int(r["# Samples"])*100/max(tot,1):5.1f}%  exec {r["Instructions Executed"]:>9s}  {why[6:]:10s} {r["Source"].strip()[:80]}')
    ex = sorted(rows, key=lambda r: -int(r['L1 Wavefronts Shared Excessive'] or 0))[:6]
    for r in ex:
        if int(r['L1 Wavefronts Shared Excessive'] or 0) > 0:
            print(f'   smem excess {r["L1 Wavefronts Shared Excessive"]:>9s} of {r["L1 Wavefronts Shared"]:>9s}  {r["Source"].strip()[:80]}')
