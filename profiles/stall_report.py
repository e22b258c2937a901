#!/usr/bin/env python
"""Per-instruction stall report of the hottest loop of a kernel from an ncu report (--import-source on).

    python profiles/stall_report.py <report.ncu-rep> <kernel-regex> [n_top]
"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', f'regex:{pat}'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': []}
        blocks.append(cur)
    elif cur is not None:
        cur['rows'].append(r)
b = blocks[0]
hdr = b['rows'][0]
data = [r for r in b['rows'][1:] if len(r) == len(hdr)]
isrc, isamp, iex = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
first = {}
for i, h in enumerate(hdr):
    if h.startswith('stall_') and h not in first:
        first[h] = i
mx = max(int(r[iex]) for r in data)
hot = [r for r in data if int(r[iex]) >= mx * 0.9]
tot = sum(int(r[isamp]) for r in hot)
print(b['name'][:90])
print(f'hot loop: {len(hot)} instructions executed {mx} times; {tot} of {sum(int(r[isamp]) for r in data)} samples')
agg = {h: sum(int(r[i] or 0) for r in hot) for h, i in first.items()}
for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
    print(f'  {h:28s} {v:7d} {v / max(tot, 1) * 100:5.1f}%')
print('--- most-sampled instructions of the hot loop')
for r in sorted(hot, key=lambda r: -int(r[isamp]))[:ntop]:
    top = sorted(((int(r[i] or 0), h[6:]) for h, i in first.items()), reverse=True)[:2]
    print(f'{int(r[isamp]):5d}  {r[isrc].strip()[:72]:72s} {top}')
