"""Host-side logic of the spin sharding (mrphy.parallel) on CPU with the gloo backend, world_size 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, 'mrphy.py_b200'))
    from mrphy import mobjs, parallel
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        cube = mobjs.SpinCube((2, 5, 4, 3), torch.tensor([[10., 8., 6.]]), dtype=torch.float64)
        cube.Δf = torch.rand((2, 5, 4, 3), generator=g, dtype=torch.float64)
        cube.T1 = 1 + torch.rand((2, 5, 4, 3), generator=g, dtype=torch.float64)
        b1 = torch.rand((2, 60, 2), generator=g, dtype=torch.float64)
        sp, kw = parallel.shard_spins(cube, rank, world, b1Map_=b1)
        lo, hi = parallel.shard_range(60, rank, world)
        ok = sp.shape == (2, hi - lo) and sp.nM == hi - lo
        ok &= torch.equal(kw['loc_'], cube.loc_[:, lo:hi]) and torch.equal(kw['Δf_'], cube.Δf_[:, lo:hi])
        ok &= torch.equal(kw['b1Map_'], b1[:, lo:hi]) and torch.equal(sp.T1_, cube.T1_[:, lo:hi])
        ok &= sp.T2_.shape == (2, hi - lo) and torch.equal(sp.M_, cube.M_[:, lo:hi])
        # one flat all-reduce sums rf.grad, gr.grad and extras over ranks, in place
        rf = torch.zeros(2, 2, 7, dtype=torch.float64, requires_grad=True)
        gr = torch.zeros(2, 3, 7, dtype=torch.float64, requires_grad=True)
        rf.grad = torch.full_like(rf, float(rank + 1))
        gr.grad = torch.arange(42, dtype=torch.float64).reshape(2, 3, 7) * (rank + 1)
        loss = torch.tensor([10.0 * (rank + 1)], dtype=torch.float64)
        parallel.allreduce_waveform_grads(rf, gr, loss)
        tot = sum(r + 1 for r in range(world))
        ok &= bool((rf.grad == tot).all()) and torch.equal(gr.grad, torch.arange(42.).reshape(2, 3, 7).double() * tot)
        ok &= float(loss) == 10.0 * tot
        # the batch axis as the shard axis: entries [lo, hi) of spins and pulse, shared (size-1) quantities untouched
        pulse = mobjs.Pulse(rf=torch.rand((2, 2, 7), generator=g, dtype=torch.float64), gr=torch.rand((2, 3, 7), generator=g,
                            dtype=torch.float64), dt=torch.tensor([4e-6, 8e-6], dtype=torch.float64), dtype=torch.float64)
        spb, kwb, pb, (blo, bhi) = parallel.shard_batch(cube, pulse, rank, world, b1Map_=b1)
        ok &= (blo, bhi) == parallel.shard_range(2, rank, world) and spb.shape == (bhi - blo, 60)
        ok &= torch.equal(pb.rf, pulse.rf[blo:bhi]) and torch.equal(pb.gr, pulse.gr[blo:bhi]) and torch.equal(pb.dt, pulse.dt[blo:bhi])
        ok &= torch.equal(kwb['loc_'], cube.loc_[blo:bhi]) and torch.equal(kwb['Δf_'], cube.Δf_[blo:bhi])
        ok &= torch.equal(kwb['b1Map_'], b1[blo:bhi]) and torch.equal(spb.T1_, cube.T1_[blo:bhi]) and torch.equal(spb.M_, cube.M_[blo:bhi])
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_shard_range_is_a_balanced_partition():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, 'mrphy.py_b200'))
    from mrphy import parallel
    for nM, world in ((16777216, 8), (10, 3), (7, 8), (262144, 1)):
        cuts = [parallel.shard_range(nM, r, world) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == nM
        assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
        sizes = [b - a for a, b in cuts]
        assert max(sizes) - min(sizes) <= 1


def test_sharding_and_allreduce_world2():
    world, port = 2, _free_port()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}
