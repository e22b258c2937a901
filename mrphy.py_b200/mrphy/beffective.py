r"""B-effective related functions (``/root/reference/mrphy/beffective.py`` surface), CUDA only.

Every function here runs a hand-written kernel of ``libmrphy_b200.so`` and carries its own explicit adjoint
(``torch.autograd.Function``); CPU tensors raise -- there is no second implementation.  ``rfgr2beff``
materialises the dense field `(N,*Nd,nT,xyz)`; the fused simulation path (``mobjs.SpinArray.applypulse`` ->
``_ops.fused_applypulse``) never calls it -- there the field is formed in registers inside the kernel.  It serves
the explicit-``Beff`` API (``Pulse.beff``, ``pulse2beff``, ``sims.blochsim(Mi, Beff)``).
"""
import math
from typing import Optional, Tuple

import torch
from torch import tensor, Tensor

from mrphy import γH, dt0, π
from mrphy import _ops

__all__ = ['beff2ab', 'beff2uφ', 'rfgr2beff']


def _reduce_to(g: Tensor, shape, tail: int = 0) -> Tensor:
    """Sum a full `(N,*Nd,<tail dims>)` gradient down to a broadcastable input of `shape`."""
    lead = g.ndim - tail
    padded = tuple(shape[:len(shape) - tail]) + (1,) * (lead - (len(shape) - tail)) + tuple(shape[len(shape) - tail:])
    return g.sum_to_size(padded).reshape(shape)


def _working(x: Tensor, like: Tensor) -> Tensor:
    """The arithmetic type of a kernel is that of its main operand; a mismatching tensor operand is cast."""
    return x if x.dtype == like.dtype else x.to(like.dtype)


class _Beff2UPhi(torch.autograd.Function):
    r"""``U = beff/max(|beff|, 1e-12)``, ``Φ = -|beff|·γ2πdt`` in one pass (28 B/spin fp32) with the adjoint of
    exactly those two expressions; upstream composes ``F.normalize`` and ``torch.norm`` (beffective.py:35-36)."""

    @staticmethod
    def forward(ctx, beff, g):
        N, Nd = beff.shape[0], tuple(beff.shape[1:-1])
        b = _ops._inner_contig(beff.reshape(N, -1, 3), 1)
        gf = _ops.flat_param(g, N, Nd, beff.device)
        U, Phi = _ops.beff2uphi_cuda(b, gf)
        ctx.save_for_backward(b, gf)
        ctx.shapes = (beff.shape, g.shape)
        return U.reshape(beff.shape), Phi.reshape(beff.shape[:-1])

    @staticmethod
    def backward(ctx, gU, gPhi):
        b, gf = ctx.saved_tensors
        N, nM = b.shape[0], b.shape[1]
        gU = None if gU is None else gU.reshape(N, nM, 3).contiguous()
        gPhi = None if gPhi is None else gPhi.reshape(N, nM).contiguous()
        gb, gg = _ops.beff2uphi_bwd_cuda(gU, gPhi, b, gf)
        bshape, gshape = ctx.shapes
        out_g = _reduce_to(gg.reshape(bshape[:-1]), gshape) if ctx.needs_input_grad[1] else None
        return gb.reshape(bshape), out_g


def beff2uϕ(beff: Tensor, γ2πdt: Tensor, *, dim=-1) -> Tuple[Tensor, Tensor]:
    r"""Rotation axes/angles from B-effective (beffective.py:18-37).

    - ``beff``: `(N, *Nd, xyz)` Gauss;  ``γ2πdt``: `()` ⊻ `(N ⊻ 1, *Nd ⊻ 1,)` rad/Gauss
    - ``dim``: position of the `xyz` axis when it is not the last one
    - returns ``U`` `(N, *Nd, xyz)` unit axes (0 where the field vanishes), ``Φ`` `(N, *Nd)` angles,
      negative because the Bloch equation is M×B.
    """
    _ops._require_cuda(beff)
    last = beff.ndim - 1
    dim = dim % beff.ndim
    b = beff if dim == last else beff.movedim(dim, last)
    if b.ndim == 1:
        b = b[None]
    U, Φ = _Beff2UPhi.apply(b, _working(_ops.on_device(γ2πdt, beff.device), beff))
    if beff.ndim == 1:
        U, Φ = U[0], Φ[0]
    return (U if dim == last else U.movedim(last, dim)), Φ


class _Beff2AB(torch.autograd.Function):
    r"""[A|B] propagated in registers by one kernel (four columns through the same rotation + relaxation step as
    the simulation) and its adjoint kernel: time-reversed columns, resynchronised every K steps with checkpoints
    of [A|B], gradients w.r.t. ``beff`` and -- reduced here onto their broadcast shapes -- E1, E2, γ, dt.
    Upstream differentiates its Python time loop by autograd (beffective.py:88-103)."""

    @staticmethod
    def forward(ctx, beff, E1, E2, γ, dt):
        N, Nd, dev = beff.shape[0], tuple(beff.shape[1:-2]), beff.device
        b = _ops._inner_contig(beff.reshape(N, -1, beff.shape[-2], 3), 2)
        E1f, E2f, γf = (_ops.flat_param(x, N, Nd, dev) for x in (E1, E2, γ))
        dtf = _ops.on_device(dt, dev).reshape(-1)
        flags = _ops.default_flags()
        K = 0
        if any(ctx.needs_input_grad):
            # un-relaxing K steps amplifies rounding by E^-K: keep that below e^0.4 (one host read of min E);
            # K = 1 is the division-free mode of the kernel, also right for E = 0
            Emin = float(torch.minimum(E1f.min(), E2f.min()))
            K = 1 if not (Emin > 0) else (_ops.K_MAX if Emin >= 1 else
                                          int(max(1, min(_ops.K_MAX, _ops._AMPLIFY_BUDGET / -math.log(Emin)))))
        A, B, ckpt = _ops.beff2ab_cuda(b, E1f, E2f, γf, dtf, K, flags)
        ctx.save_for_backward(A, B, ckpt, b, E1f, E2f, γf, dtf)
        ctx.K, ctx.flags = K, flags
        ctx.shapes = (beff.shape, E1.shape, E2.shape, γ.shape, dt.shape)
        lead = beff.shape[:-2]
        return A.reshape(lead + (3, 3)), B.reshape(lead + (3,))

    @staticmethod
    def backward(ctx, gA, gB):
        A, B, ckpt, b, E1f, E2f, γf, dtf = ctx.saved_tensors
        N, nM = b.shape[0], b.shape[1]
        zero = lambda ref: torch.zeros_like(ref)
        gA = zero(A) if gA is None else gA.reshape(N, nM, 3, 3).contiguous()
        gB = zero(B) if gB is None else gB.reshape(N, nM, 3).contiguous()
        gb, gP = _ops.beff2ab_bwd_cuda(gA, gB, A, B, ckpt, b, E1f, E2f, γf, dtf, ctx.K, ctx.flags)
        bshape, s1, s2, sγ, sdt = ctx.shapes
        lead, need = bshape[:-2], ctx.needs_input_grad
        full = lambda x: x.reshape(lead)
        out = [gb.reshape(bshape) if need[0] else None, None, None, None, None]
        if need[1]:
            out[1] = _reduce_to(full(gP[..., 0]), s1)
        if need[2]:
            out[2] = _reduce_to(full(gP[..., 1]), s2)
        if need[3] or need[4]:       # g = 2π·γ·dt per spin
            gg = gP[..., 2]
            dtb = dtf.reshape(-1, 1).to(gg.dtype)
            if need[3]:
                out[3] = _reduce_to(full(gg * (2 * π) * dtb), sγ)
            if need[4]:
                out[4] = _reduce_to(full(gg * (2 * π) * γf.to(gg.dtype)), sdt)
        return tuple(out)


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
    _ops._require_cuda(beff)
    dev = beff.device
    mv = lambda x: _ops.on_device(x, dev)
    return _Beff2AB.apply(beff, mv(E1), mv(E2), mv(γ), mv(dt))


class _RfGr2Beff(torch.autograd.Function):
    r"""CUDA field synthesis (one pass, 12 B/spin·step written) with the chain rule of beffective.py:137-167
    written out: the reference gets these gradients from autograd through bmm / broadcast / stack.  Backward: the sums
    over spins (∂rf, ∂gr) and the sums over time (∂loc, ∂Δf, ∂b1Map, ∂γ) are ONE kernel each over the dense ∂L/∂Beff."""

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
            b1n = b1Map if b1Map.ndim == len(Nd) + 3 else b1Map[..., None]          # (N|1,*Nd|1,2,nC|1)
            # a single-coil b1Map broadcasts over the coils of rf, as upstream (beffective.py:153-165); the opposite
            # case (multi-coil b1Map, rf without coils) is reduced to one coil by `rfgr2beff` below
            assert b1n.shape[-1] in (1, nC), 'b1Map and rf disagree on nCoils'
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
        gB = gB.reshape(N, nM, -1, 3).contiguous()
        out = [None] * 6
        if need[0] or need[1]:       # the two sums over spins: one native pass over dL/dBeff
            grf, ggr = _ops.rfgr2beff_bwd_cuda(gB, rf, gr, loc, b1)
            out[0], out[1] = (grf if need[0] else None), (ggr if need[1] else None)
        want_loc, want_b1 = bool(need[2]), bool(need[4] and b1 is not None)
        want_z = df is not None and (need[3] or need[5])
        if want_loc or want_b1 or want_z:    # the sums over time (per-spin gradients): one more native pass
            gloc, gsz, gb1 = _ops.rfgr2beff_spin_grads_cuda(gB, rf, gr, want_loc, want_b1)
            if want_loc:
                out[2] = gloc.reshape(ctx.shapes[0])
            if want_z:
                gzs = gsz.reshape((N,) + Nd)
                γf = γ.expand(N, nM).reshape((N,) + Nd)
                if need[3]:
                    out[3] = _reduce_to(gzs / γf, ctx.shapes[1])
                if need[5]:
                    dff = df.expand(N, nM).reshape((N,) + Nd)
                    out[5] = _reduce_to(-gzs * dff / (γf * γf), ctx.shapes[3])
            if want_b1:
                gb = gb1.reshape((N,) + Nd + (2, -1))                                      # (N,*Nd,2,nC)
                shp = ctx.shapes[2]
                if len(shp) == len(Nd) + 2:                                                # b1Map given without coil dim
                    out[4] = _reduce_to(gb[..., 0], shp, tail=1)
                else:
                    out[4] = _reduce_to(gb, shp, tail=2)
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
    _ops._require_cuda(rf)
    dev = rf.device
    cast = lambda x: None if x is None else _working(x.to(dev), rf)
    b1Map = cast(b1Map)
    if b1Map is not None and b1Map.ndim == loc.ndim + 1 and b1Map.shape[-1] > 1 and (rf.ndim == 3 or rf.shape[3] == 1):
        # the same rf drives every coil: sum_c b1_c * rf = (sum_c b1_c) * rf  (upstream's broadcast, beffective.py:158-165)
        b1Map = b1Map.sum(dim=-1, keepdim=True)
    return _RfGr2Beff.apply(rf, cast(gr), cast(loc), cast(Δf), b1Map, _ops.on_device(γ, dev))
