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
from mrphy import utils

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
