r"""Autograd-differentiated simulators (``blochsim_1step``, ``blochsim``, ``blochsim_ab``, ``freeprec``).

The reference ships these (``/root/reference/mrphy/slowsims.py``) as its own cross-check of the hand-written
Jacobians in ``sims``; they are kept here for API completeness only.  They are a few self-contained,
device-agnostic torch expressions inside a Python loop over time, differentiable w.r.t. every argument by plain
autograd -- and are NOT the product path: nothing in ``sims``, ``beffective`` or ``mobjs`` comes through this
module, and this module calls none of the CUDA operators.
"""
from typing import Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

from mrphy import dt0, utils, γH, π

__all__ = ['blochsim_1step', 'blochsim', 'blochsim_ab', 'freeprec']

_Opt = Optional[Tensor]


def _advance(M: Tensor, b: Tensor, rad_per_gauss: Tensor, E1: Tensor, E2: Tensor) -> Tensor:
    """One dwell time: precess about ``b`` (skipped where the whole field is zero), then T2 decay / T1 recovery.
    ``E1`` is `(N,*Nd)`-broadcastable, ``E2`` already carries a trailing singleton for the xy pair."""
    axis, angle = F.normalize(b, dim=-1), -torch.norm(b, dim=-1) * rad_per_gauss   # M×B: negative angle
    if torch.any(angle != 0):
        M = utils.uϕrot(axis, angle, M)
    xy = M[..., 0:2] * E2
    z = M[..., 2:3] * E1[..., None] + (1 - E1)[..., None]
    return torch.cat((xy, z), dim=-1)


def blochsim_1step(M: Tensor, M1: Tensor, b: Tensor, E1: Tensor, E1_1: Tensor, E2: Tensor,
                   γ2πdt: Tensor) -> Tuple[Tensor, Tensor]:
    r"""Advance ``M`` `(N,*Nd,xyz)` by one step in the field ``b`` `(N,*Nd,xyz)` [Gauss] (slowsims.py:15-57).

    ``E1``, ``E2`` are the per-step relaxation factors, ``E1_1 = E1 - 1`` and ``M1`` are accepted for signature
    compatibility (upstream uses them as scratch).  Returns ``(M_new, M_old)``.
    """
    return _advance(M, b, γ2πdt, E1, E2[..., None]), M


def blochsim(M: Tensor, Beff: Tensor, *, T1: _Opt = None, T2: _Opt = None, γ: Tensor = γH, dt: Tensor = dt0) -> Tensor:
    r"""``Mo = blochsim(M, Beff, *, T1, T2, γ, dt)`` with the arguments of ``sims.blochsim`` (slowsims.py:60-114)."""
    assert M.shape[:-1] == Beff.shape[:-2]
    dev, lead = M.device, M.ndim - 1
    one = torch.ones((), device=dev, dtype=M.dtype)
    E1 = one if T1 is None else torch.exp(-dt / T1.to(dev))
    E2 = one if T2 is None else torch.exp(-dt / T2.to(dev))
    E1, E2, γ, dt = (utils._tail(t.to(dev), lead) for t in (E1, E2, γ, dt))
    rad_per_gauss = 2 * π * γ * dt
    Beff = Beff.to(dev)
    for t in range(Beff.shape[-2]):
        M = _advance(M, Beff[..., t, :], rad_per_gauss, E1, E2[..., None])
    return M


def blochsim_ab(M: Tensor, A: Tensor, B: Tensor) -> Tensor:
    r"""Apply Hargreaves' propagator: ``A @ M + B`` per spin, ``A`` `(N,*Nd,xyz,3)`, ``B`` `(N,*Nd,xyz)`."""
    return torch.matmul(A, M.unsqueeze(-1)).squeeze(-1) + B


def freeprec(M: Tensor, dur: Tensor, *, T1: _Opt = None, T2: _Opt = None, Δf: _Opt = None) -> Tensor:
    r"""``M = freeprec(M, dur, *, T1, T2, Δf)`` with the arguments of ``sims.freeprec`` (slowsims.py:134-174)."""
    rank = M.ndim
    dur = utils._tail(dur, rank)
    x, y, z = M.unbind(dim=-1)
    x, y, z = x[..., None], y[..., None], z[..., None]
    if Δf is not None:       # a positive off-resonance turns the spin clockwise
        turn = -(2 * π) * utils._tail(Δf, rank) * dur
        c, s = turn.cos(), turn.sin()
        x, y = c * x - s * y, s * x + c * y
    assert (T1 is None) == (T2 is None)
    if T1 is not None:
        E1 = torch.exp(-dur / utils._tail(T1, rank))
        E2 = torch.exp(-dur / utils._tail(T2, rank))
        x, y, z = E2 * x, E2 * y, E1 * z + (1 - E1)
    return torch.cat((x, y, z), dim=-1)
