r"""B-effective related functions (``/root/reference/mrphy/beffective.py`` surface).

``rfgr2beff`` materialises the dense field `(N,*Nd,nT,xyz)`; the fused simulation path
(``mobjs.SpinArray.applypulse`` -> ``_ops.fused_applypulse``) never calls it -- the field is formed in
registers inside the CUDA kernel.  It is kept for the explicit-``Beff`` API (``Pulse.beff``,
``pulse2beff``, ``sims.blochsim(Mi, Beff)``) and stays differentiable w.r.t. every input by plain
autograd, like upstream.
"""
from typing import Optional, Tuple

import torch
import torch.nn.functional as F
from torch import tensor, Tensor

from mrphy import γH, dt0, π
from mrphy import utils, _ops

__all__ = ['beff2ab', 'beff2uφ', 'rfgr2beff']


def beff2uϕ(beff: Tensor, γ2πdt: Tensor, *, dim=-1) -> Tuple[Tensor, Tensor]:
    r"""Rotation axes/angles from B-effective (beffective.py:18-37).

    - ``beff``: `(N, *Nd, xyz)` Gauss;  ``γ2πdt``: `()` ⊻ `(N ⊻ 1, *Nd ⊻ 1,)` rad/Gauss
    - returns ``U`` `(N, *Nd, xyz)` unit axes (0 where the field vanishes), ``Φ`` `(N, *Nd)` angles,
      negative because the Bloch equation is M×B.
    """
    return F.normalize(beff, dim=dim), -torch.norm(beff, dim=dim) * γ2πdt


def beff2ab(
    beff: Tensor, *,
    E1: Tensor = tensor(0.), E2: Tensor = tensor(0.),
    γ: Tensor = γH, dt: Tensor = dt0,
) -> Tuple[Tensor, Tensor]:
    r"""Hargreaves 𝐴/𝐵 (doi:10.1002/mrm.1170) from B-effective (beffective.py:40-104).

    - ``beff``: `(N,*Nd,nT,xyz)`;  ``E1``, ``E2``: per-step relaxation factors exp(-dt/T), `()` ⊻
      `(N ⊻ 1, *Nd ⊻ 1,)` (NB: factors, not times -- same as upstream despite its docstring)
    - returns ``A`` `(N,*Nd,xyz,3)`, ``B`` `(N,*Nd,xyz)` with ``M_end = A @ M_start + B``.
    """
    dev, nd = beff.device, beff.ndim - 2
    if beff.is_cuda and beff.dtype in _ops._F and not (torch.is_grad_enabled() and beff.requires_grad):
        # forward-only CUDA kernel (one spin per thread, [A|B] in registers); gradients use the loop below
        N, Nd = beff.shape[0], tuple(beff.shape[1:-2])
        flat = lambda x: _ops.flat_param(x, N, Nd, dev)
        b = _ops._inner_contig(beff.reshape(N, -1, beff.shape[-2], 3), 2)
        A, B = _ops.beff2ab_cuda(b, flat(E1), flat(E2), flat(γ), dt.to(dev).reshape(-1), _ops.default_flags())
        return A.reshape(beff.shape[:-2] + (3, 3)), B.reshape(beff.shape[:-2] + (3,))
    E1, E2, γ, dt = (utils._tail(x.to(dev), nd) for x in (E1, E2, γ, dt))
    g = 2 * π * γ * dt
    NNd, nT = beff.shape[:-2], beff.shape[-2]
    AB = torch.eye(3, 4, device=dev, dtype=beff.dtype).expand(NNd + (3, 4)).clone()   # [A | B]
    scale = torch.stack((E2, E2, E1), dim=-1)[..., None].to(beff.dtype)               # (N,*Nd,3,1) rows
    recover = (1 - E1).to(beff.dtype)
    for t in range(nT):
        u, ϕ = beff2uϕ(beff[..., t, :], g)
        AB = utils.uϕrot(u, ϕ, AB) if torch.any(ϕ != 0) else AB
        AB = AB * scale
        AB = torch.cat((AB[..., :3], AB[..., 3:] + torch.stack(
            (torch.zeros_like(recover), torch.zeros_like(recover), recover), dim=-1)[..., None].expand(NNd + (3, 1))),
            dim=-1)
    return AB[..., 0:3], AB[..., 3]


def _reduce_to(g: Tensor, shape, tail: int = 0) -> Tensor:
    """Sum a full `(N,*Nd,<tail dims>)` gradient down to a broadcastable input of `shape`."""
    lead = g.ndim - tail
    padded = tuple(shape[:len(shape) - tail]) + (1,) * (lead - (len(shape) - tail)) + tuple(shape[len(shape) - tail:])
    return g.sum_to_size(padded).reshape(shape)


class _RfGr2Beff(torch.autograd.Function):
    r"""CUDA field synthesis (one pass, 12 B/spin·step written) with the chain rule of beffective.py:137-167
    written out: the reference gets these gradients from autograd through bmm / broadcast / stack."""

    @staticmethod
    def forward(ctx, rf, gr, loc, Δf, b1Map, γ):
        N, Nd = loc.shape[0], tuple(loc.shape[1:-1])
        dev = rf.device
        locf = _ops._inner_contig(loc.reshape(N, -1, 3), 1)
        nM = locf.shape[1]
        dff = None if Δf is None else _ops.flat_param(Δf, N, Nd, dev)
        gf = _ops.flat_param(γ, N, Nd, dev)
        b1f = None
        if b1Map is not None:
            nC = rf.shape[3] if rf.ndim == 4 else 1
            b1n = b1Map if b1Map.ndim == len(Nd) + 3 else b1Map[..., None]          # (N|1,*Nd|1,2,nC)
            assert b1n.shape[-1] == nC, 'b1Map and rf disagree on nCoils'
            b1f = _ops._inner_contig(b1n.expand((N,) + Nd + (2, nC)).reshape(N, nM, 2, nC), 2)
        out = _ops.rfgr2beff_cuda(rf, gr, locf, dff, b1f, gf)
        ctx.save_for_backward(rf, gr, locf, dff, b1f, gf)
        ctx.Nd = Nd
        ctx.shapes = (loc.shape, None if Δf is None else Δf.shape, None if b1Map is None else b1Map.shape, γ.shape)
        return out.reshape((N,) + Nd + (rf.shape[2], 3))

    @staticmethod
    def backward(ctx, gB):
        rf, gr, loc, df, b1, γ = ctx.saved_tensors
        need, Nd = ctx.needs_input_grad, ctx.Nd
        N, nM = loc.shape[0], loc.shape[1]
        gB = gB.reshape(N, nM, -1, 3)
        gx, gy, gz = gB.unbind(-1)
        out = [None] * 6
        rf4 = rf if rf.ndim == 4 else rf[..., None]
        if need[0]:
            if b1 is None:
                g2 = torch.stack((gx.sum(1), gy.sum(1)), dim=1)                       # (N,2,nT)
                out[0] = g2[..., None].expand(rf4.shape).contiguous() if rf.ndim == 4 else g2
            else:
                br, bi = b1[:, :, 0], b1[:, :, 1]                                      # (N,nM,nC)
                grx = torch.einsum('nmc,nmt->ntc', br, gx) + torch.einsum('nmc,nmt->ntc', bi, gy)
                gry = torch.einsum('nmc,nmt->ntc', br, gy) - torch.einsum('nmc,nmt->ntc', bi, gx)
                g4 = torch.stack((grx, gry), dim=1)
                out[0] = g4 if rf.ndim == 4 else g4[..., 0]
        if need[1]:
            out[1] = torch.einsum('nmc,nmt->nct', loc, gz)
        if need[2]:
            out[2] = torch.einsum('nct,nmt->nmc', gr, gz).reshape(ctx.shapes[0])
        gzs = gz.sum(-1).reshape((N,) + Nd) if (df is not None and (need[3] or need[5])) else None
        γf = γ.expand(N, nM).reshape((N,) + Nd) if gzs is not None else None
        if need[3] and df is not None:
            out[3] = _reduce_to(gzs / γf, ctx.shapes[1])
        if need[4] and b1 is not None:
            rx, ry = rf4[:, 0], rf4[:, 1]                                              # (N,nT,nC)
            gbr = torch.einsum('ntc,nmt->nmc', rx, gx) + torch.einsum('ntc,nmt->nmc', ry, gy)
            gbi = torch.einsum('ntc,nmt->nmc', rx, gy) - torch.einsum('ntc,nmt->nmc', ry, gx)
            gb = torch.stack((gbr, gbi), dim=2).reshape((N,) + Nd + (2, -1))          # (N,*Nd,2,nC)
            shp = ctx.shapes[2]
            if len(shp) == len(Nd) + 2:                                                # b1Map given without coil dim
                out[4] = _reduce_to(gb[..., 0], shp, tail=1)
            else:
                out[4] = _reduce_to(gb, shp, tail=2)
        if need[5] and df is not None:
            dff = df.expand(N, nM).reshape((N,) + Nd)
            out[5] = _reduce_to(-gzs * dff / (γf * γf), ctx.shapes[3])
        return tuple(out)


def rfgr2beff(
    rf: Tensor, gr: Tensor, loc: Tensor, *,
    Δf: Optional[Tensor] = None, b1Map: Optional[Tensor] = None, γ: Tensor = γH
) -> Tensor:
    r"""B-effective from rf and gradients (beffective.py:107-168).

    - ``rf``: `(N,xy,nT,(nCoils))` Gauss;  ``gr``: `(N,xyz,nT)` Gauss/cm;  ``loc``: `(N,*Nd,xyz)` cm
    - ``Δf``: `(N,*Nd,)` Hz;  ``b1Map``: `(N,*Nd,xy,(nCoils))`;  ``γ``: `()` ⊻ `(N ⊻ 1, *Nd ⊻ 1,)` Hz/Gauss
    - returns ``beff``: `(N,*Nd,nT,xyz)` Gauss
    """
    assert (rf.device == gr.device == loc.device)
    dev = rf.device
    N, Nd = loc.shape[0], tuple(loc.shape[1:-1])
    nd = len(Nd)
    same = all(x is None or (x.dtype == rf.dtype) for x in (gr, loc, Δf, b1Map))
    if rf.is_cuda and rf.dtype in _ops._F and same and γ.dtype in _ops._F:
        return _RfGr2Beff.apply(rf, gr, loc, Δf, None if b1Map is None else b1Map.to(dev), γ.to(dev))
    Bz = torch.matmul(loc.reshape(N, -1, 3), gr).reshape((N,) + Nd + (-1,))        # loc·gr
    if Δf is not None:
        Bz = Bz + utils._tail(Δf, nd + 2) / utils._tail(γ.to(device=dev), nd + 2)   # off-resonance as a z-field
    rfx = rf.reshape((N,) + nd * (1,) + tuple(rf.shape[1:]))                         # (N,1..,xy,nT,(nCoils))
    if b1Map is None:
        if rfx.ndim == Bz.ndim + 2:
            rfx = rfx.sum(dim=-1)                                                     # coils add up
        Bx, By = rfx[..., 0, :].expand_as(Bz), rfx[..., 1, :].expand_as(Bz)
    else:
        b1 = b1Map.to(dev)
        if b1.ndim == nd + 2:
            b1 = b1[..., None]
        if rfx.ndim == b1.ndim:
            rfx = rfx[..., None]
        br, bi = b1[..., 0, None, :], b1[..., 1, None, :]                            # (N,*Nd,1,nCoils)
        rx, ry = rfx[..., 0, :, :], rfx[..., 1, :, :]                                # (N,1..,nT,nCoils)
        Bx = (br * rx - bi * ry).sum(dim=-1).expand_as(Bz)                           # Re(b1·rf)
        By = (br * ry + bi * rx).sum(dim=-1).expand_as(Bz)                           # Im(b1·rf)
    return torch.stack((Bx, By, Bz), dim=-1)
