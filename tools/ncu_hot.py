#!/usr/bin/env python
"""Top stalled SASS instructions from `ncu -i X.ncu-rep --page source --csv` (needs -lineinfo builds).
usage: ncu_hot.py source.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot = sum(f(r, '# Samples') for r in data)
inst = sum(f(r, 'Instructions Executed') for r in data)
print(f'total samples {tot:.0f}, warp instructions executed {inst:.0f}, sass lines {len(data)}')
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: sum(f(r, s) for r in data) for s in stalls}
print('stall mix:', ', '.join(f'{k[6:]} {v / max(tot,1) * 100:.1f}%' for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.005 * tot))
print('--- top by samples')
for r in sorted(data, key=lambda r: -f(r, '# Samples'))[:n]:
    top = sorted(((f(r, s), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"{r[ix['Address']][-5:]} {f(r, '# Samples') / tot * 100:5.2f}%  exec {f(r, 'Instructions Executed'):12.0f}  {top[0][1]:12s} {top[1][1]:12s} {r[ix['Source']][:90]}")
