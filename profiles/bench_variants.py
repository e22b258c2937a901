#!/usr/bin/env python
"""Kernel-only comparison of tuning variants of the fused kernels (C2 / C3, fp32), one subprocess per variant.

    python profiles/bench_variants.py name[:lib.so][:ENV=V,ENV=V] ...  [--workloads c2,c3] [--trig precise,mixed,fast]

Each variant is a library built with `python mrphy.py_b200/build.py -D... --out=profiles/variants/x.so` and/or a set of
MRPHY_B200_* environment switches.  Prints per variant: forward / backward kernel ms (CUDA events inside the C ABI, best and
median of 8 after 3 warm-ups, 256-MB L2 flush between steps), spin-steps/s, and max |grad - grad(first variant)| so a
variant that is fast because it is wrong shows up."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker(workload, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'mrphy.py_b200'))
    import numpy as np
    import torch
    import bench
    from mrphy import mobjs, _cabi
    dev = torch.device('cuda:0')
    kw = {'dtype': torch.float32, 'device': dev}
    N, n, nT = bench.WORKLOADS[workload]
    d = {k: v.to(dev) for k, v in bench.synth(N, n, n, nT, torch.float32, interp=5 if workload == 'c3' else 1).items()}
    sp = mobjs.SpinArray((N, d['loc'].shape[1]), M_=d['M0'], **kw)
    pulse = mobjs.Pulse(rf=d['rf'].requires_grad_(True), gr=d['gr'].requires_grad_(True), **kw)
    tgt = torch.tensor([0., 1., 0.], **kw)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    L = _cabi.lib()
    L.mrphy_kernel_timing(1)
    f, b = [], []
    for i in range(11):
        flush.fill_(1)
        pulse.rf.grad = pulse.gr.grad = None
        M = sp.applypulse(pulse, loc_=d['loc'], Δf_=d['df'], b1Map_=d['b1'])
        tf = L.mrphy_last_kernel_ms()
        ((M - tgt) ** 2).sum().backward()
        tb = L.mrphy_last_kernel_ms()
        if i >= 3:
            f.append(tf)
            b.append(tb)
    torch.cuda.synchronize()
    units = float(N) * d['loc'].shape[1] * nT
    res = {'fwd_ms': min(f), 'bwd_ms': min(b), 'fwd_med': float(np.median(f)), 'bwd_med': float(np.median(b)),
           'rate_best': units / ((min(f) + min(b)) * 1e-3), 'rate_med': units / ((np.median(f) + np.median(b)) * 1e-3)}
    np.savez(out_path, grf=pulse.rf.grad.cpu().numpy(), ggr=pulse.gr.grad.cpu().numpy(), M=M.detach()[:, ::97].cpu().numpy())
    print(json.dumps(res))


def main():
    import numpy as np
    specs = [a for a in sys.argv[1:] if not a.startswith('--')]
    opts = dict(a[2:].split('=', 1) for a in sys.argv[1:] if a.startswith('--'))
    workloads = opts.get('workloads', 'c2').split(',')
    trigs = opts.get('trig', 'precise').split(',')
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    for wl in workloads:
        for trig in trigs:
            ref = None
            for spec in specs:
                parts = spec.split(':')
                name = parts[0]
                env = dict(os.environ)
                if len(parts) > 1 and parts[1]:
                    env['MRPHY_B200_LIB'] = os.path.join(ROOT, parts[1])
                if len(parts) > 2 and parts[2]:
                    env.update(dict(kv.split('=', 1) for kv in parts[2].split(',')))
                env['MRPHY_B200_TRIG'] = trig
                outp = os.path.join(ROOT, 'gpurun_out', f'variant_{name}_{wl}_{trig}.npz')
                r = subprocess.run([sys.executable, os.path.abspath(__file__), '--worker', wl, outp], env=env,
                                   capture_output=True, text=True)
                if r.returncode != 0:
                    print(f'{wl} {trig} {name}: FAILED\n{r.stderr[-1500:]}')
                    continue
                res = json.loads(r.stdout.strip().splitlines()[-1])
                z = np.load(outp)
                if ref is None:
                    ref = z
                sc = lambda k: float(np.abs(z[k] - ref[k]).max() / (np.abs(ref[k]).max() + 1e-30))
                print(f"{wl} {trig:7s} {name:16s} fwd {res['fwd_ms']:.4f} (med {res['fwd_med']:.4f})  bwd {res['bwd_ms']:.4f} "
                      f"(med {res['bwd_med']:.4f})  {res['rate_med']:.4e} spin-steps/s  "
                      f"rel diff vs first: grf {sc('grf'):.2e} ggr {sc('ggr'):.2e} M {sc('M'):.2e}", flush=True)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == '--worker':
        worker(sys.argv[2], sys.argv[3])
    else:
        main()
