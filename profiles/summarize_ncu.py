#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into profiles/r2_ncu_summary.json entries (MRPHY_NCU_SUMMARY overrides the file).

    python profiles/summarize_ncu.py gpurun_out/prof_final.ncu-rep final     # -> captures.final_fwd / final_bwd / ...
    python profiles/summarize_ncu.py --launches profiles/r1_launches_c2.csv   # -> launch_list_c2_default (shares)

Runs here (no GPU): `ncu -i ... --page raw --csv` only reads the report.
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.environ.get('MRPHY_NCU_SUMMARY') or os.path.join(ROOT, 'profiles', 'r2_ncu_summary.json')
WANT = [
    'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
    'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__cycles_elapsed.avg.per_second', 'sm__cycles_active.avg',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__warps_active.avg.per_cycle_active', 'smsp__warps_eligible.avg.per_cycle_active',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
]
TAGS = (('fused_bwd', 'bwd'), ('fused_fwd', 'fwd'), ('grad_finalize', 'finalize'), ('pack_waveform', 'pack'),
        ('rfgr2beff_bwd', 'rfgr2beff_bwd'), ('rfgr2beff', 'rfgr2beff'), ('beff_v3_kernel<float, 1, 1, 1>', 'beff_bwd'),
        ('beff_v3_kernel<float, 1, 1, 0>', 'beff_fwd'), ('beff_v3', 'beff'), ('beff_v2', 'beff'))


def num(x):
    return float(x.replace(',', ''))


def captures(rep, prefix):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index('Kernel Name')
    out = {}
    for r in rows[2:]:
        tag = next((t for k, t in TAGS if k in r[ki]), None)
        if tag is None:
            continue
        m = {}
        for w in WANT:
            if w in hdr:
                try:
                    m[w] = {'value': num(r[hdr.index(w)]), 'unit': units[hdr.index(w)]}
                except ValueError:
                    pass
        out[f'{prefix}_{tag}'] = {'kernel': r[ki], 'metrics': m}     # the last launch of a kernel wins
    return out


def launch_shares(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows[1:]:
        v = num(r[vi]) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r[ui], 1.0)
        name = r[ki].split('(')[0]
        name = name if 'mrphy' in name else 'torch: ' + name.split('<')[0].replace('void ', '')
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    return {k: {'launches': cnt[k], 'us': round(v, 1), 'share_pct': round(100 * v / T, 2)} for k, v in sorted(tot.items(), key=lambda x: -x[1])}


if __name__ == '__main__':
    d = json.load(open(OUT)) if os.path.exists(OUT) else {'captures': {}}
    if sys.argv[1] == '--launches':
        d[sys.argv[3] if len(sys.argv) > 3 else 'launch_list_c2_default'] = launch_shares(sys.argv[2])
    else:
        d.setdefault('captures', {}).update(captures(sys.argv[1], sys.argv[2]))
    json.dump(d, open(OUT, 'w'), indent=1)
    print('updated', OUT)
