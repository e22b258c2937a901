#!/usr/bin/env python
"""Offline summary of the default fp32 kernels of a library build: registers / stack and the register-operand-bandwidth
model (profiles/operand_model.py) of their 4-step main loops.  No GPU needed.

    python profiles/model_all.py [lib.so ...]
"""
import re
import subprocess
import sys

sys.path.insert(0, __import__('os').path.dirname(__file__))
from operand_model import model, parse_full  # noqa: E402
from sass_census import kernels, loops  # noqa: E402

KERNELS = {   # label: (mangled-name substring, steps per main-loop iteration)
    'fwd precise': 'fused_fwd_kernelIfLi1ELb1ELi1ELi2ELi64E',
    'bwd precise': 'fused_bwd_kernelIfLi1ELb1ELi1ELi2ELi128ELi3E',
    'bwd fast': 'fused_bwd_kernelIfLi0ELb1ELi1ELi2ELi128ELi3E',
    'fwd fast': 'fused_fwd_kernelIfLi0ELb1ELi1ELi2ELi64E',
}


def res_usage(lib):
    out = subprocess.run(['cuobjdump', '-res-usage', lib], capture_output=True, text=True).stdout
    res, cur = {}, None
    for l in out.splitlines():
        m = re.search(r'Function (\S+):', l)
        if m:
            cur = m.group(1)
        elif cur and 'REG:' in l:
            res[cur] = ' '.join(re.findall(r'(?:REG|STACK|SHARED):\d+', l))
            cur = None
    return res


def main_loop(lines):
    """the innermost loop with >= 200 instructions (the 4-step body)"""
    ins = parse_full(lines)
    best = None
    for lo, hi in loops(ins):
        body = [(a, op, rest) for a, op, rest in ins if lo <= a <= hi]
        if len(body) >= 200 and (best is None or len(body) < len(best)):
            best = body
    return best


if __name__ == '__main__':
    for lib in sys.argv[1:] or ['mrphy.py_b200/libmrphy_b200.so']:
        ks, ru = kernels(lib), res_usage(lib)
        print(f'== {lib}')
        tot = {}
        for label, sub in KERNELS.items():
            for name, lines in ks.items():
                if sub not in name:
                    continue
                body = main_loop(lines)
                if body is None:
                    continue
                issue, pipe, operands, cyc, reuse, rd = model(body)
                tot[label] = cyc / 4
                print(f'  {label:12s} {ru.get(name, "?"):28s} issue {issue:4d} fma-pipe {pipe:4d} reads {rd:4d} reuse {reuse:3d} '
                      f'-> model {cyc / 4:6.1f} cycles/thread-step')
        if 'fwd precise' in tot and 'bwd fast' in tot:
            for mode, b in (('mixed', 'bwd fast'), ('precise', 'bwd precise')):
                s = tot['fwd precise'] + tot[b] + 25.0
                print(f'  {mode:8s}: fwd + bwd + 25 (spin reduction) = {s:6.1f} cycles/thread-step; nominal issue roofline 302 -> '
                      f'{302 / s:.3f} x (measured/model ~0.91-0.95)')
