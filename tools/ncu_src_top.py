"""Top stall sites per kernel from `ncu -i X.ncu-rep --page source --csv` (SASS view).
usage: ncu -i rep --page source --csv [--kernel-name regex:NAME] | python tools/ncu_src_top.py [N] [name-filter]"""
import csv
import sys

n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
flt = sys.argv[2] if len(sys.argv) > 2 else ''
rows = [r for r in csv.reader(sys.stdin)]
blocks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'hdr': None, 'body': []}
        blocks.append(cur)
    elif cur is not None and cur['hdr'] is None and 'Source' in r:
        cur['hdr'] = r
    elif cur is not None and cur['hdr'] is not None and len(r) == len(cur['hdr']):
        cur['body'].append(r)
for b in blocks:
    if flt and flt not in b['name']:
        continue
    hdr, body = b['hdr'], b['body']
    ia, isamp = hdr.index('Source'), hdr.index('# Samples')
    stall = [x for x in hdr if x.startswith('stall_') and 'Not Issued' not in x]
    cols = {x: hdr.index(x) for x in stall}
    tot = sum(int(r[isamp]) for r in body)
    print('==', b['name'][:80], ' total samples', tot)
    agg = {k: sum(int(r[c]) for r in body) for k, c in cols.items()}
    print({k.replace('stall_', ''): f'{100 * v / max(tot, 1):.1f}%' for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
    for r in sorted(body, key=lambda r: -int(r[isamp]))[:n]:
        print(f'{100 * int(r[isamp]) / max(tot, 1):5.1f}%', r[ia][:100], {k.replace('stall_', ''): int(r[c]) for k, c in cols.items() if int(r[c]) > 0})
