#!/usr/bin/env python
"""Static register-operand-bandwidth model of a kernel's loops (cuobjdump -sass, no GPU needed).

    python profiles/operand_model.py <mangled-kernel-substring> [lib.so] [--min-instr 100]

Measured on B200 (profiles/ubench_ffma2_operands.txt): an SM sub-partition reads about TWO 32-bit register operands per
lane per cycle -- FFMA with three distinct registers 1.5 cycles, FMUL 1.0, FFMA2 with three register pairs 3.0, with two
pairs + immediate 2.0, and operands served by the reuse cache (`.reuse`) or given as immediates / constants are free.
Per loop this prints
    issue     = instructions (1 issue slot each)
    pipe      = FMA-pipe cycles (packed FFMA2/FMUL2/FADD2 count 2)
    operands  = 0.5 * register source operands that are not `.reuse` hits
    model     = sum over instructions of max(1, pipe_i, 0.5 * reads_i)   -- cycles per loop iteration per warp
so `model` / (steps per iteration) is the predicted cycles per thread-step to compare with
    measured = kernel time * clock * 148 SMs * 4 * 32 lanes / (threads * steps)."""
import re
import subprocess
import sys

from sass_census import kernels, loops

PACKED = ('FFMA2', 'FMUL2', 'FADD2')
NO_DEST = ('STS', 'STG', 'ST', 'RED', 'REDG', 'BAR', 'BRA', 'EXIT', 'SYNCS', 'UBLKCP', 'ATOMS', 'NOP', 'WARPSYNC', 'BSYNC',
           'BSSY', 'ISETP', 'FSETP', 'PLOP3', 'R2UR')
WIDTH = {'64': 2, '128': 4}


def parse_full(lines):
    ins = []
    for l in lines:
        m = re.search(r'/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)\s*(.*?);', l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2), m.group(3)))
    return ins


def reads(op, rest):
    """number of 32-bit vector-register source operands that go through the register file"""
    base = op.split('.')[0]
    ops = [o.strip() for o in rest.split(',')] if rest else []
    if base in ('ISETP', 'FSETP'):
        srcs = ops[2:] if len(ops) > 2 else ops     # two predicate destinations
    elif base in NO_DEST:
        srcs = ops
    else:
        srcs = ops[1:]
    n = 0
    wide = next((WIDTH[w] for w in WIDTH if ('.' + w) in op), 1)
    for i, o in enumerate(srcs):
        for r in re.finditer(r'\bR(\d+)((?:\.\w+)*)', o):
            mods = r.group(2)
            if '.reuse' in mods:
                continue
            k = 1
            if base in PACKED and 'F32x2' in mods:
                k = 2
            elif base in ('STS', 'STG', 'ST', 'REDG') and '[' not in o:
                k = wide                      # data registers of a wide store
            elif '.64' in mods:
                k = 2
            n += k
    return n


def model(body):
    issue = len(body)
    pipe = sum(2 if op.split('.')[0] in PACKED else 1 for _, op, _ in body
               if op.split('.')[0] in PACKED + ('FFMA', 'FMUL', 'FADD', 'IMAD', 'HFMA2'))
    rd = [reads(op, rest) for _, op, rest in body]
    operands = 0.5 * sum(rd)
    tot = sum(max(1.0, 2.0 if op.split('.')[0] in PACKED else 1.0, 0.5 * r) for (_, op, _), r in zip(body, rd))
    reuse = sum(rest.count('.reuse') for _, _, rest in body)
    return issue, pipe, operands, tot, reuse, sum(rd)


if __name__ == '__main__':
    args = [a for a in sys.argv[1:] if not a.startswith('--')]
    lib = args[1] if len(args) > 1 else 'mrphy.py_b200/libmrphy_b200.so'
    min_instr = int(next((a.split('=')[1] for a in sys.argv if a.startswith('--min-instr=')), 100))
    for name, lines in kernels(lib).items():
        if args[0] not in name:
            continue
        ins = parse_full(lines)
        print(f'== {name}')
        for lo, hi in sorted(loops([(a, op, rest) for a, op, rest in ins]), key=lambda x: x[1] - x[0]):
            body = [(a, op, rest) for a, op, rest in ins if lo <= a <= hi]
            if len(body) < min_instr:
                continue
            issue, pipe, operands, tot, reuse, rd = model(body)
            print(f'  loop {lo:#06x}-{hi:#06x}: issue {issue:5d}  fma-pipe {pipe:5d}  operand-cycles {operands:7.1f} '
                  f'(reads {rd}, reuse hits {reuse})  model {tot:7.1f} cycles/iteration')
