#!/usr/bin/env python
"""One iteration of a slew- and amplitude-constrained pulse design: unconstrained variables (tρ, θ, ts) ->
rf = tρθ2rf, gr = s2g(ts2s) -> cube.applypulse -> loss -> backward to the variables.  Compares the re-parametrisation as
the reference's torch expressions (MRPHY_B200_REPARAM=torch), as one kernel per utils function, as the single fused
launch `utils.tρθts2rfgr`, and with the chain's adjoint folded into the simulation's gradient epilogue ("+tail", the default;
the other rows run with MRPHY_B200_FUSE_DESIGN=0); reports ms per iteration (CUDA events, eager launches) and kernels launched
per iteration."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'mrphy.py_b200'))
import torch
import bench
from mrphy import mobjs, utils, rfmax0, smax0

dev = torch.device('cuda:0'); kw = {'dtype': torch.float32, 'device': dev}
from mrphy import dt0
rfmax0, smax0, dt0 = rfmax0.to(**kw), smax0.to(**kw), dt0.to(**kw)   # the torch expressions need them on the device


def iteration(mode, sp, d, v, tgt):
    tρ, θ, ts = v
    if mode.startswith('fused'):
        rf, gr = utils.tρθts2rfgr(tρ, θ, ts, rfmax0, smax0, dt0)
    else:
        rf = utils.tρθ2rf(tρ, θ, rfmax0)
        gr = utils.s2g(utils.ts2s(ts, smax0), dt0)
    pulse = mobjs.Pulse(rf=rf, gr=gr, **kw)
    M = sp.applypulse(pulse, loc_=d['loc'], Δf_=d['df'], b1Map_=d['b1'])
    loss = ((M - tgt) ** 2).sum()
    for x in v:
        x.grad = None
    loss.backward()
    return loss


def count_kernels(fn):
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    return sum(e.count for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA and 'Memcpy' not in e.key and 'Memset' not in e.key)


for n, nT in ((16, 256), (32, 512), (64, 1000)):
    d = {k: t.to(dev) for k, t in bench.synth(1, n, n, nT, torch.float32).items()}
    sp = mobjs.SpinArray((1, d['loc'].shape[1]), M_=d['M0'], **kw)
    tgt = torch.tensor([0., 1., 0.], **kw)
    g = torch.Generator(device='cuda').manual_seed(1)
    v = [torch.randn((1, c, nT), generator=g, **kw).mul_(0.3).requires_grad_(True) for c in (1, 1, 3)]
    ref = None
    for mode in ('torch', 'per-function', 'fused', 'fused+tail', 'fused+tail+graph'):
        os.environ['MRPHY_B200_REPARAM'] = 'torch' if mode == 'torch' else 'cuda'
        os.environ['MRPHY_B200_FUSE_DESIGN'] = '1' if 'tail' in mode else '0'
        for _ in range(5):
            iteration(mode, sp, d, v, tgt)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 50 if n < 64 else 20
        run = lambda: iteration(mode, sp, d, v, tgt)
        if mode.endswith('+graph'):           # the whole iteration as one CUDA graph (it is free of host synchronisation)
            from mrphy import graphs
            loss = None
            cap = graphs.capture(run, params=v)
            run = cap.replay
        e0.record()
        for _ in range(K):
            loss = run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        grads = torch.cat([x.grad.reshape(-1) for x in v])
        if ref is None:
            ref = grads.clone()
        try:
            nk = count_kernels(run)
        except Exception as ex:   # profiler unavailable: report time only
            nk = f'n/a ({type(ex).__name__})'
        print(f'{n}^3 x {nT}: {mode:17s} {ms:7.3f} ms/iteration  kernels/iteration {nk}  loss {float(loss):.6e}  '
              f'max rel grad diff vs torch chain {float((grads - ref).abs().max() / ref.abs().max()):.2e}', flush=True)
