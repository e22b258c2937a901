#!/usr/bin/env python
"""bench.py -- spin·steps/s of the fused Bloch simulation, forward + adjoint backward.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4|c5|small]
                    [--dtype f32|f64]

One "step" = one pass of the hot path over one batch of synthetic input (BASELINE.md / SURVEY.md 8d):
``M = cube.applypulse(pulse, b1Map_=...)``; ``loss = ((M - target)**2).sum()``; ``loss.backward()`` giving
``pulse.rf.grad``, ``pulse.gr.grad`` (+ one NCCL all-reduce of the waveform gradient when N > 1).
Default workload at N=1 is BASELINE config C2: SpinCube 64^3 (262 144 spins), nT=1000, dt=4us, fp32.
With N GPUs every rank owns one such slab of a (64*N) x 64 x 64 cube (weak scaling; waveform replicated).

Prints ONE JSON line (rank 0).  ``value`` is measured with inputs resident in HBM (CUDA events around each
step, L2 flushed between steps; the step is replayed from a CUDA graph unless --no-graph); ``e2e`` goes through the
same public API eagerly but copies every input from pinned host memory every step and reads the loss and gradients
back into pinned memory, all inside one wall-clock timed region of K steps (uploads of step i+1 overlap step i).
``--impl reference`` times the CPU port of the reference's algorithm (oracle/bloch_oracle.py: torch CPU ops,
one handful per time step like the reference) on the host cores -- the reference itself is Python and cannot
travel to the GPU box.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, 'mrphy.py_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = 'spin_steps_per_sec_fwd_bwd'
UNIT = 'spin·steps/s'
FLOP_PER_SPIN_STEP = 223.0        # fwd 62 + bwd 161 algorithmic flop (fp32, 1 coil, relax, b1, df)
FLOP_FWD, FLOP_BWD = 62.0, 161.0
WORKLOADS = {   # name: (N, n, nT)
    'c2': (1, 64, 1000), 'c3': (1, 128, 2000), 'c4': (64, 40, 1000), 'c5': (1, 256, 4000), 'small': (1, 16, 200),
}


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            mp = json.load(f)
        return float(mp['hbm_gbs']), float(mp['sm_max_mhz']), 'measured'
    except Exception:
        return 6650.0, 1965.0, 'fallback'


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace('.', '').isdigit()]
        mxs = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower() == 'active'})
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mxs) if mxs else None,
                'reasons': reasons, 'samples': len(sm)}


def synth(N, n_x, n, nT, dtype, seed=0, x_off=0, n_x_total=None, interp=1):
    """Seeded synthetic slab (SURVEY 8d distributions), generated in fp64 then rounded to `dtype`.
    Slab = x-indices [x_off, x_off+n_x) of a (n_x_total, n, n) grid with 24 cm fov per 64 voxels.
    interp > 1 (BASELINE config C3, "multi-scale interpT design"): the waveform is drawn with nT/interp samples at
    interp*4 us in fp64 and brought to nT samples at 4 us by `Pulse.interpT` (fp64, or the grid is 1 short)."""
    n_x_total = n_x if n_x_total is None else n_x_total
    gen = torch.Generator().manual_seed(seed + 1000 * x_off)
    U = lambda *s: torch.rand(s, generator=gen, dtype=torch.float64) * 2 - 1
    fov = torch.tensor([24.0 * n_x_total / n, 24.0, 24.0], dtype=torch.float64)
    ax = [(torch.arange(x_off, x_off + n_x, dtype=torch.float64) - n_x_total // 2) / n_x_total,
          (torch.arange(n, dtype=torch.float64) - n // 2) / n, (torch.arange(n, dtype=torch.float64) - n // 2) / n]
    g = torch.meshgrid(*ax, indexing='ij')
    loc = torch.stack([fov[i] * g[i].reshape(-1) for i in range(3)], dim=-1)[None].expand(N, -1, 3).contiguous()
    nM = loc.shape[1]
    wgen = torch.Generator().manual_seed(seed)          # the waveform is the same on every rank
    W = lambda *s: torch.rand(s, generator=wgen, dtype=torch.float64) * 2 - 1
    if interp > 1:
        from mrphy import mobjs
        assert nT % interp == 0
        f8 = torch.float64
        coarse = mobjs.Pulse(rf=W(N, 2, nT // interp) * 0.1, gr=W(N, 3, nT // interp) * 2,
                             dt=torch.tensor(4e-6 * interp, dtype=f8), dtype=f8)
        fine = coarse.interpT(dt=torch.tensor(4e-6, dtype=f8))
        rf, gr = fine.rf, fine.gr
        assert rf.shape[2] == nT, rf.shape
    else:
        rf, gr = W(N, 2, nT) * 0.1, W(N, 3, nT) * 2
    df = U(N, nM) * 200
    b1 = U(N, nM, 2) * 0.1
    b1[:, :, 0] += 1
    M0 = torch.tensor([0., 0., 1.], dtype=torch.float64).expand(N, nM, 3).contiguous()
    return {k: v.to(dtype) for k, v in dict(rf=rf, gr=gr, loc=loc, df=df, b1=b1, M0=M0).items()}


def workload_desc(args, world):
    N, n, nT = WORKLOADS[args.workload]
    return (f'{args.workload.upper()}: SpinCube {n}^3 x N={N} ' +
            ('per GPU' if args.scaling == 'weak' else f'split over {world} GPUs') +
            f', nT={nT}, dt=4us' + (f' (Pulse.interpT from {nT // 5} x 20us)' if args.workload == 'c3' else '') +
            ', b1Map+df+relaxation, fwd+adjoint bwd')


def run_ours(args):
    from mrphy import mobjs, parallel, _cabi
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        # rank 0 prints ONE JSON line on stdout: NCCL's version banner / debug output goes to stderr instead
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        if os.environ.get('NCCL_DEBUG', '').upper() in ('VERSION', 'WARN'):
            del os.environ['NCCL_DEBUG']          # the banner ignores NCCL_DEBUG_FILE (checked: profiles/nccl_stdout_check.py)
        dist.init_process_group('nccl', device_id=dev)
    N, n, nT = WORKLOADS[args.workload]
    dtype = torch.float32 if args.dtype == 'f32' else torch.float64
    kw = {'dtype': dtype, 'device': dev}
    itp = 5 if args.workload == 'c3' else 1    # C3: the pulse comes out of Pulse.interpT (400 x 20 us -> 2000 x 4 us)
    if args.scaling == 'strong':      # the whole n^3 cube split into x-slabs (BASELINE config C5)
        assert n % world == 0
        host = synth(N, n // world, n, nT, dtype, x_off=rank * (n // world), n_x_total=n, interp=itp)
    else:                             # weak: every rank owns an n^3 slab of an (n*world) x n x n cube
        host = synth(N, n, n, nT, dtype, x_off=rank * n, n_x_total=n * world, interp=itp)
    pinned = {k: v.pin_memory() for k, v in host.items()}
    nM = host['loc'].shape[1]
    tgt = torch.tensor([0., 1., 0.], **kw)

    def make_objects(src, non_blocking=False):
        d = {k: v.to(dev, non_blocking=non_blocking) for k, v in src.items()}
        sp = mobjs.SpinArray((N, nM), M_=d['M0'], **kw)
        pulse = mobjs.Pulse(rf=d['rf'].requires_grad_(True), gr=d['gr'].requires_grad_(True), **kw)
        return sp, pulse, d

    def local_step(sp, pulse, d):
        M = sp.applypulse(pulse, loc_=d['loc'], Δf_=d['df'], b1Map_=d['b1'])
        loss = ((M - tgt) ** 2).sum()
        loss.backward()
        return loss

    def step(sp, pulse, d):
        loss = local_step(sp, pulse, d)
        if world > 1:
            parallel.allreduce_waveform_grads(pulse.rf, pulse.gr, loss.detach().reshape(1))
        return loss

    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # 256 MB > 126 MB L2
    sp, pulse, d = make_objects(host)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)      # samples through warm-up, the timed region and the e2e / per-kernel legs
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        pulse.rf.grad = pulse.gr.grad = None
        step(sp, pulse, d)
    barrier()
    # ---- timed region: resident inputs, per-step CUDA events, L2 flushed between steps.  The local part of a step
    # (forward, loss, adjoint backward: 4 of our kernels + the loss kernels) is recorded once into a CUDA graph and
    # replayed; the gradient all-reduce (N > 1) stays eager.  --no-graph, or a failed capture, falls back to eager
    # launches of the same calls.
    def timed_region():
        graph, g_loss, per_step = None, None, None
        if not args.no_graph:
            try:
                pulse.rf.grad = pulse.gr.grad = None
                l0 = _cabi.launch_counter
                g_ = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_):
                    g_loss = local_step(sp, pulse, d)
                per_step = _cabi.launch_counter - l0
                for _ in range(2):
                    g_.replay()
                graph = g_
            except Exception as e:                  # noqa: BLE001 -- any capture problem: measure eagerly instead
                print(f'[bench] CUDA graph capture failed ({type(e).__name__}: {e}); timing eager launches',
                      file=sys.stderr)
                torch.cuda.synchronize()
                graph = None

        def one():
            if graph is None:
                pulse.rf.grad = pulse.gr.grad = None
                step(sp, pulse, d)
            else:
                graph.replay()
                if world > 1:
                    parallel.allreduce_waveform_grads(pulse.rf, pulse.gr, g_loss.detach().reshape(1))

        barrier()
        l0 = _cabi.launch_counter
        evs = []
        for _ in range(args.steps):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            one()
            e1.record()
            evs.append((e0, e1))
        barrier()
        ms = sum(x.elapsed_time(y) for x, y in evs)
        n_launch = per_step * args.steps if graph is not None else _cabi.launch_counter - l0
        mode = 'eager' if graph is None else 'CUDA graph replay of the step (pack, fwd, loss, bwd, finalize)'
        pulse.rf.grad = pulse.gr.grad = None
        return ms, n_launch, mode

    ms_total, launches, launch_mode = timed_region()
    # ---- end-to-end: pinned host inputs -> device every step, loss + gradients read back into pinned host memory.
    # The objects live across steps as in a design loop; every step overwrites ALL their device data from pinned host
    # memory.  Two object sets: the copy stream uploads step i+1 while the compute stream runs step i (what any input
    # pipeline does); the timed region is the whole K-step loop, wall clock, synchronised on both sides, so every
    # upload and every read-back is inside it.  `serial_ms_per_step` is the un-pipelined latency of one such step.
    sets = [make_objects(host), make_objects(host)]
    outs = [(torch.empty(1, dtype=dtype).pin_memory(), torch.empty_like(host['rf']).pin_memory(),
             torch.empty_like(host['gr']).pin_memory()) for _ in range(2)]
    copy_stream = torch.cuda.Stream()
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    free = [torch.cuda.Event(), torch.cuda.Event()]

    def upload(b, stream):
        sp_b, pulse_b, d_b = sets[b]
        with torch.cuda.stream(stream), torch.no_grad():
            for k in ('loc', 'df', 'b1'):
                d_b[k].copy_(pinned[k], non_blocking=True)
            sp_b.M_.copy_(pinned['M0'], non_blocking=True)
            pulse_b.rf.copy_(pinned['rf'], non_blocking=True)
            pulse_b.gr.copy_(pinned['gr'], non_blocking=True)

    def compute(b):
        sp_b, pulse_b, d_b = sets[b]
        pulse_b.rf.grad = pulse_b.gr.grad = None
        loss = step(sp_b, pulse_b, d_b)
        o = outs[b]
        o[0].copy_(loss.detach().reshape(1), non_blocking=True)      # results: three async copies into pinned memory
        o[1].copy_(pulse_b.rf.grad, non_blocking=True)
        o[2].copy_(pulse_b.gr.grad, non_blocking=True)
        return o

    def e2e_serial():
        upload(0, torch.cuda.current_stream())
        o = compute(0)
        torch.cuda.current_stream().synchronize()
        return o

    def e2e_pipelined(K):
        cur = torch.cuda.current_stream()
        for ev in free:
            ev.record(cur)
        copy_stream.wait_stream(cur)
        upload(0, copy_stream)
        ready[0].record(copy_stream)
        for i in range(K):
            b = i & 1
            if i + 1 < K:                       # next step's inputs, once the set they overwrite is no longer in use
                copy_stream.wait_event(free[b ^ 1])
                upload(b ^ 1, copy_stream)
                ready[b ^ 1].record(copy_stream)
            flush.fill_(1.0)                    # inside the timed region here (cannot be hidden in a pipeline)
            cur.wait_event(ready[b])
            o = compute(b)
            free[b].record(cur)
        torch.cuda.synchronize()
        return o

    for _ in range(2):
        e2e_serial()
    e2e_pipelined(3)
    barrier()
    t_serial = []
    for _ in range(min(args.steps, 5)):
        flush.fill_(1.0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_serial()
        t_serial.append(time.perf_counter() - t0)
    barrier()
    t0 = time.perf_counter()
    out = e2e_pipelined(args.steps)
    t_e2e = [time.perf_counter() - t0]
    barrier()
    h2d = sum(v.numel() * v.element_size() for v in pinned.values())
    d2h = sum(o.numel() * o.element_size() for o in out)
    del sets, outs
    # ---- per-kernel durations (events inside the C ABI, on the launching stream)
    L = _cabi.lib()
    L.mrphy_kernel_timing(1)
    k_fwd, k_bwd = [], []
    for _ in range(min(args.steps, 10)):
        flush.fill_(1.0)
        pulse.rf.grad = pulse.gr.grad = None
        M = sp.applypulse(pulse, loc_=d['loc'], Δf_=d['df'], b1Map_=d['b1'])
        k_fwd.append(L.mrphy_last_kernel_ms())
        ((M - tgt) ** 2).sum().backward()
        k_bwd.append(L.mrphy_last_kernel_ms())
    L.mrphy_kernel_timing(0)
    del M          # a live autograd graph would pin AccumulateGrad nodes to this stream and spoil the next capture
    # ---- the opt-in MUFU trigonometry (MRPHY_B200_TRIG=fast), same timed region, reported as an extra
    alt_ms, mixed_ms = None, None
    if dtype == torch.float32:
        for pol in ('fast', 'mixed'):
            os.environ['MRPHY_B200_TRIG'] = pol
            for _ in range(2):
                pulse.rf.grad = pulse.gr.grad = None
                step(sp, pulse, d)
            ms = timed_region()[0]
            alt_ms, mixed_ms = (ms, mixed_ms) if pol == 'fast' else (alt_ms, ms)
        del os.environ['MRPHY_B200_TRIG']
    clocks = sampler.stop() if rank == 0 else None
    # ---- reduce over ranks (max time), aggregate
    t = torch.tensor([ms_total, sum(t_e2e) * 1e3, alt_ms or 0.0, mixed_ms or 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, alt_ms, mixed_ms = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    units = float(N) * nM * nT * world            # spin·steps per step, all ranks (nM is per rank)
    value = units * args.steps / (ms_total * 1e-3)
    e2e = units * args.steps / (ms_e2e * 1e-3)
    if rank == 0:
        hbm_gbs, sm_mhz, src = peaks()
        peak_tf = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
        ms_b, ms_f = float(np.mean(k_bwd)), float(np.mean(k_fwd))
        per_launch = float(N) * nM * nT
        ach_tf = per_launch * FLOP_BWD / (ms_b * 1e-3) / 1e12
        esz = 4 if dtype == torch.float32 else 8
        K = 64
        ck_bytes = per_launch / K * 3 * esz + N * nM * (3 + 3 + 3 + 2 + 1 + 3) * esz   # bwd: ckpt reads + operands
        traffic = None
        try:   # dram__bytes_read+write per launch of the same kernel from the committed ncu capture (C2 fp32 only)
            if args.workload == 'c2' and args.dtype == 'f32':
                with open(os.path.join(ROOT, 'profiles', 'r1_ncu_summary.json')) as f:
                    m = json.load(f)['captures']['final_bwd']['metrics']
                traffic = (m['dram__bytes_read.sum']['value'] + m['dram__bytes_write.sum']['value']) * 1e6
        except Exception:
            traffic = None
        operand = None
        try:   # the bound that actually binds these kernels: register-operand bandwidth, static SASS model
            if args.dtype == 'f32':
                with open(os.path.join(ROOT, 'profiles', 'r1_operand_model.json')) as f:
                    om = json.load(f)
                clk = sm_mhz * 1e6
                lanes = 148 * 4 * 32
                meas = lambda ms: ms * 1e-3 * clk * lanes / (per_launch / 2)        # cycles per thread-step (2 spins/thread)
                mf, mb = om['fwd']['model_cycles_per_thread_step'], om['bwd']['model_cycles_per_thread_step'] + 25.0
                operand = {'what': 'cycles per thread-step (two spins) of one SM sub-partition: static register-operand-'
                                   'bandwidth model of the SASS main loops (+25 for the backward spin reduction) vs measured',
                           'fwd': {'model': mf, 'measured': meas(ms_f), 'frac': mf / meas(ms_f)},
                           'bwd': {'model': mb, 'measured': meas(ms_b), 'frac': mb / meas(ms_b)},
                           'fwd_bwd_frac': (mf + mb) / (meas(ms_f) + meas(ms_b)),
                           'source': 'profiles/r1_operand_model.json, profiles/ubench_ffma2_operands.txt'}
        except Exception:
            operand = None
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_total / args.steps, 'higher_is_better': True,
            'scaling': args.scaling, 'vs_baseline': None, 'dtype': args.dtype, 'data': 'synthetic',
            'config': {'workload': workload_desc(args, world), 'spins_per_gpu': N * nM, 'nT': nT,
                       'l2': 'flushed between steps (256 MB write)', 'launch': launch_mode, 'sharding': f'spin slabs x{world}, waveform '
                       'replicated, 1 allreduce of grads' if world > 1 else 'single GPU'},
            'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'how': 'K steps back to back, wall clock: uploads of step i+1 (copy stream) overlap step i, L2 flush and '
                           'read-back inside the timed region', 'serial_ms_per_step': float(np.mean(t_serial)) * 1e3},
            'gpu_launches': launches,
            'clocks': clocks,
            'roofline': {'bound': 'fp32_issue', 'kernel': 'fused_bwd_kernel', 'achieved': ach_tf, 'peak': peak_tf,
                         'unit': 'TFLOP/s', 'frac': ach_tf / peak_tf, 'traffic': traffic, 'traffic_unit': 'bytes/launch',
                         'algorithmic_bytes_per_launch': ck_bytes,
                         'peak_source': f'148 SM x 128 FP32 lanes x 2 x sm_max_mhz ({src} {sm_mhz:.0f} MHz); the path '
                                        'is FP32-issue-bound (SURVEY 8d), not HBM- or tensor-bound',
                         'ms_per_launch': ms_b, 'algorithmic_flop_per_spin_step': FLOP_BWD,
                         'register_operand_bound': operand,
                         'fwd_kernel': {'ms_per_launch': ms_f, 'achieved': per_launch * FLOP_FWD / (ms_f * 1e-3) / 1e12,
                                        'frac': per_launch * FLOP_FWD / (ms_f * 1e-3) / 1e12 / peak_tf},
                         'fwd_bwd_frac_of_issue_roofline': value / world / (148 * 128 * sm_mhz * 1e6 / 151.0),
                         'hbm': {'achieved_gbs': ck_bytes / (ms_b * 1e-3) / 1e9, 'peak_gbs': hbm_gbs,
                                 'frac': ck_bytes / (ms_b * 1e-3) / 1e9 / hbm_gbs, 'source': src}},
        }
        if alt_ms:
            line['alt'] = {'trig': 'fast (MUFU.SIN/COS/RSQ, MRPHY_B200_TRIG=fast)', 'value': units * args.steps / (alt_ms * 1e-3),
                           'unit': UNIT, 'ms_per_step': alt_ms / args.steps,
                           'note': 'opt-in: ~2x the fp32 error of the default polynomial trigonometry'}
        if mixed_ms:
            line['alt_mixed'] = {'trig': 'precise forward, MUFU trigonometry in the adjoint only (MRPHY_B200_TRIG=mixed)',
                                 'value': units * args.steps / (mixed_ms * 1e-3), 'unit': UNIT,
                                 'ms_per_step': mixed_ms / args.steps,
                                 'note': 'opt-in: M identical to the default; rf/gr gradients 2-3e-5 relative instead of '
                                         '4-7e-6 (the reference\'s own fp32: ~1.5e-5; tolerance 1e-4)'}
        line['config']['trig'] = 'precise (default)' if dtype == torch.float32 else 'fp64 libm'
        if world == 1:      # CPU / eager baselines are timed at N=1 only; the other ranks must not wait on them
            line['cpu_baseline'] = cpu_baseline(args, nT)
            try:
                line['torch_eager_same_gpu'] = eager_cuda_baseline(args, nT)
            except Exception as e:      # context only
                line['torch_eager_same_gpu'] = {'error': str(e)[:100]}
        else:
            line['cpu_baseline'] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_port_step(n, nT, dtype, threads):
    """One fwd+bwd of the oracle port on an n^3 proxy cube with the bench distributions."""
    from oracle import bloch_oracle as orc
    torch.set_num_threads(threads)
    s = synth(1, n, n, nT, dtype)
    t0 = time.perf_counter()
    tgt = torch.tensor([0., 1., 0.], dtype=dtype)
    orc.applypulse_fwd_bwd(s['M0'], s['rf'], s['gr'], s['loc'], lambda Mo: 2 * (Mo - tgt), df=s['df'], b1=s['b1'],
                           T1=1.47, T2=0.07, dtype=dtype)
    return time.perf_counter() - t0, n ** 3 * nT


def eager_cuda_baseline(args, nT):
    """The reference's execution shape (a handful of torch ops per time step, dense Beff and per-step history in
    HBM) run on THIS GPU: the oracle port with its tensors on cuda.  Context only -- not the product, not the target."""
    from oracle import bloch_oracle as orc
    dtype = torch.float32 if args.dtype == 'f32' else torch.float64
    n = 32
    s = {k: v.cuda() for k, v in synth(1, n, n, nT, dtype).items()}
    tgt = torch.tensor([0., 1., 0.], dtype=dtype, device='cuda')
    orc.DEVICE = 'cuda'
    try:
        best = None
        for _ in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            orc.applypulse_fwd_bwd(s['M0'], s['rf'], s['gr'], s['loc'], lambda Mo: 2 * (Mo - tgt), df=s['df'], b1=s['b1'],
                                   T1=1.47, T2=0.07, dtype=dtype)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    finally:
        orc.DEVICE = 'cpu'
    return {'value': n ** 3 * nT / best, 'unit': UNIT, 'kind': 'port of the reference algorithm in torch-eager CUDA on this GPU',
            'sample': f'{n}^3 spins x {nT} steps, {best:.2f} s (about 60 kernel launches per time step)'}


def cpu_baseline(args, nT):
    cores = len(os.sched_getaffinity(0))
    dtype = torch.float32 if args.dtype == 'f32' else torch.float64
    n = 16
    sec, units = cpu_port_step(n, nT, dtype, cores)
    return {'value': units / sec, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': f'oracle/bloch_oracle.py (torch CPU, {cores} threads) on a {n}^3 proxy cube, same nT={nT}, '
                      f'{args.dtype}, 1 fwd+bwd pass, {sec:.1f} s; the reference needs 52 B/spin-step so the full '
                      'cube does not fit / finish (BASELINE.md sec. 4)'}


def run_reference(args):
    """The reference's CPU implementation of the path, restated (oracle port), all host threads."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    N, n, nT = WORKLOADS[args.workload]
    dtype = torch.float32 if args.dtype == 'f32' else torch.float64
    cores = len(os.sched_getaffinity(0))
    budget = min(8.0, 150.0 / max(args.steps + args.warmup, 1))       # seconds per step
    n_cpu = min(n, int(max(8, min(32, round((budget * 7e5 / nT) ** (1 / 3))))))
    for _ in range(args.warmup):
        cpu_port_step(n_cpu, nT, dtype, cores)
    tot, units = 0.0, 0
    for _ in range(args.steps):
        s, u = cpu_port_step(n_cpu, nT, dtype, cores)
        tot += s
        units += u
    v = units / tot
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': int(os.environ.get('WORLD_SIZE', '1')),
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': tot / args.steps * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': args.dtype, 'data': 'synthetic',
        'config': {'workload': workload_desc(args, int(os.environ.get('WORLD_SIZE', '1'))), 'nT': nT,
                   'sample': f'each step = one fwd+bwd over a {n_cpu}^3 sub-cube of the workload (same distributions, same '
                             f'nT); spin·steps/s is size-insensitive at fixed nT (BASELINE.md sec. 2) and the full cube '
                             'needs 52 B/spin-step in the reference'},
        'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': f'{n_cpu}^3 spins x {nT} steps per step, torch CPU {cores} threads'},
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }), flush=True)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c2', choices=sorted(WORKLOADS))
    ap.add_argument('--dtype', default='f32', choices=['f32', 'f64'])
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'])
    ap.add_argument('--no-graph', action='store_true', help='time eager launches instead of a CUDA graph replay')
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == 'ours' else a.warmup
    (run_ours if a.impl == 'ours' else run_reference)(a)
