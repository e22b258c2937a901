#!/usr/bin/env python
"""Static SASS evidence per kernel of libmrphy_b200.so -> profiles/sass_mnemonics.txt (cuobjdump -sass, no GPU needed).

    python profiles/sass_mnemonics.py [lib.so] > profiles/sass_mnemonics.txt

UBLKCP = 1-D TMA bulk copy (cp.async.bulk), SYNCS = mbarrier ops, LDGSTS = cp.async 16 B, FFMA2/FMUL2/FADD2 = packed fp32,
DFMA = FP64 pipe, MUFU = XU, VOTE/CALL = the large-angle guard of the half-angle coefficients, REDG/RED = partial-sum RED,
UTCHMMA = tcgen05.mma (kind::tf32), LDTM = tcgen05.ld (TMEM -> registers), UTCBAR = tcgen05.commit, UTCATOMSWS = TMEM alloc."""
import collections
import re
import subprocess
import sys

WANT = ('UTCHMMA', 'LDTM', 'UTCBAR', 'UTCATOMSWS', 'UBLKCP', 'SYNCS', 'LDGSTS', 'FFMA2', 'FMUL2', 'FADD2', 'FFMA', 'DFMA', 'MUFU', 'VOTE', 'CALL', 'REDG', 'RED', 'SHFL', 'LDS', 'STS')

lib = sys.argv[1] if len(sys.argv) > 1 else 'mrphy.py_b200/libmrphy_b200.so'
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
cur, cnt = None, {}
for l in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', l)
    if m:
        cur = m.group(1)
        cnt[cur] = collections.Counter()
        continue
    m = re.search(r'/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)', l)
    if cur and m:
        cnt[cur][m.group(1)] += 1
names = subprocess.run(['c++filt'], input='\n'.join(cnt), capture_output=True, text=True).stdout.splitlines()
print('# SASS mnemonic counts per kernel of libmrphy_b200.so (cuobjdump -sass; static counts; profiles/sass_mnemonics.py).')
print('# ' + __doc__.split('\n\n')[2].replace('\n', '\n# '))
print()
for mangled, name in sorted(zip(cnt, names), key=lambda x: x[1]):
    name = re.sub(r'^void mrphy::', '', name)
    name = re.sub(r'\(.*$', '', name).replace('(int)', '').replace('(bool)', '').replace('(mrphy::TrigPolicy)', '')
    c = cnt[mangled]
    print(f'{name[:84]:84s} ' + ' '.join(f'{k}={c[k]}' for k in WANT if c[k]))
