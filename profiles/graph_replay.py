#!/usr/bin/env python
"""Launch-bound end of the path: a small design step (SpinCube n^3, nT steps) issued eagerly vs replayed from a CUDA
graph (the whole applypulse forward + loss + adjoint backward is capturable, tests/test_gpu_parity.py)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'mrphy.py_b200'))
import torch
from torch import tensor
from mrphy import mobjs

dev = torch.device('cuda:0'); kw = {'dtype': torch.float32, 'device': dev}
out = {}
for n, nT in ((8, 128), (16, 256), (32, 512), (64, 1000)):
    gen = torch.Generator().manual_seed(0)
    cube = mobjs.SpinCube((1, n, n, n), tensor([[24., 24., 24.]]), **kw)
    cube.Δf = (torch.rand(1, n, n, n, generator=gen) * 200 - 100).to(dev)
    pulse = mobjs.Pulse(rf=((torch.rand(1, 2, nT, generator=gen) * 0.2 - 0.1).to(dev)).requires_grad_(True),
                        gr=((torch.rand(1, 3, nT, generator=gen) * 4 - 2).to(dev)).requires_grad_(True), **kw)
    sp, loc, df = cube.spinarray, cube.loc_, cube.Δf_
    tgt = tensor([0., 1., 0.], **kw)

    def step():
        M = sp.applypulse(pulse, loc_=loc, Δf_=df)
        loss = ((M - tgt) ** 2).sum()
        loss.backward()
        return loss

    def timed(fn, reps=50):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3

    def eager():
        pulse.rf.grad = pulse.gr.grad = None
        step()

    t_eager = timed(eager)
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            eager()
    torch.cuda.current_stream().wait_stream(side)
    pulse.rf.grad = pulse.gr.grad = None
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    t_graph = timed(g.replay)
    out[f'{n}^3 x {nT}'] = {'eager_ms': round(t_eager, 4), 'graph_ms': round(t_graph, 4),
                            'spin_steps_per_s_graph': n ** 3 * nT / (t_graph * 1e-3)}
print(json.dumps(out))
