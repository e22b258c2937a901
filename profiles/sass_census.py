#!/usr/bin/env python
"""Static SASS census of the loops of one kernel in libmrphy_b200.so (cuobjdump -sass).

    python profiles/sass_census.py <mangled-kernel-substring> [lib.so]

Prints every backward-branch loop with its instruction mix, so the per-step issue-slot count can be
compared with the algorithmic budget in DESIGN.md."""
import collections
import re
import subprocess
import sys


def kernels(lib):
    out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
    cur, body = None, {}
    for l in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', l)
        if m:
            cur = m.group(1)
            body[cur] = []
        elif cur and re.search(r'/\*[0-9a-f]{4,}\*/', l):
            body[cur].append(l)
    return body


def parse(lines):
    ins = []
    for l in lines:
        m = re.search(r'/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)(.*?);', l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2), m.group(3)))
    return ins


def loops(ins):
    res = []
    for a, op, rest in ins:
        if op.startswith('BRA'):
            m = re.search(r'0x([0-9a-f]+)', rest)
            if m and int(m.group(1), 16) <= a:
                res.append((int(m.group(1), 16), a))
    return res


if __name__ == '__main__':
    lib = sys.argv[2] if len(sys.argv) > 2 else 'mrphy.py_b200/libmrphy_b200.so'
    for name, lines in kernels(lib).items():
        if sys.argv[1] not in name:
            continue
        ins = parse(lines)
        print(f'== {name}: {len(ins)} instructions')
        for lo, hi in sorted(loops(ins), key=lambda x: x[1] - x[0]):
            body = [(a, op) for a, op, _ in ins if lo <= a <= hi]
            c = collections.Counter(op.split('.')[0] for _, op in body)
            fp = sum(c[k] for k in ('FFMA', 'FMUL', 'FADD', 'FFMA2', 'FMUL2', 'FADD2'))
            print(f'  loop {lo:#06x}-{hi:#06x}: {len(body):4d} instr, FP32-pipe {fp}, MUFU {c["MUFU"]}, '
                  f'LDS {c["LDS"]}, STS {c["STS"]}, SHFL {c["SHFL"]}, BAR {c["BAR"]} | '
                  + ' '.join(f'{k}:{v}' for k, v in c.most_common(12)))
