r"""Simulation with implicit (autograd) Jacobians -- the reference's own cross-check module
(``/root/reference/mrphy/slowsims.py``), kept for API completeness.

NOT the product path: these are short device-agnostic torch expressions with a Python loop over
time, differentiable w.r.t. everything by plain autograd.  ``sims.blochsim`` /
``SpinArray.applypulse`` never route here.
"""
from typing import Optional, Tuple

import torch
from torch import tensor, Tensor

from mrphy import γH, dt0, π
from mrphy import utils, beffective

__all__ = ['blochsim_1step', 'blochsim', 'blochsim_ab', 'freeprec']


def _relax(M: Tensor, E1: Tensor, E1_1: Tensor, E2: Tensor) -> Tensor:
    return torch.cat((M[..., 0:2] * E2, M[..., 2:3] * E1[..., None] - E1_1[..., None]), dim=-1)


def blochsim_1step(M: Tensor, M1: Tensor, b: Tensor, E1: Tensor, E1_1: Tensor, E2: Tensor,
                   γ2πdt: Tensor) -> Tuple[Tensor, Tensor]:
    r"""One step: rotate ``M`` `(N,*Nd,xyz)` about ``b`` `(N,*Nd,xyz)` then relax (slowsims.py:15-57).

    ``M1`` is accepted for signature compatibility (upstream uses it as scratch); returns ``(M_new, M_old)``.
    """
    u, ϕ = beffective.beff2uϕ(b, γ2πdt)
    Mr = utils.uϕrot(u, ϕ, M) if torch.any(ϕ != 0) else M
    return _relax(Mr, E1, E1_1, E2[..., None]), M


def blochsim(M: Tensor, Beff: Tensor, *, T1: Optional[Tensor] = None, T2: Optional[Tensor] = None,
             γ: Tensor = γH, dt: Tensor = dt0) -> Tensor:
    r"""Bloch simulator with implicit Jacobians (slowsims.py:60-114); same arguments as ``sims.blochsim``."""
    assert (M.shape[:-1] == Beff.shape[:-2])
    dev, nd = M.device, M.ndim - 1
    kw = {'device': dev, 'dtype': M.dtype}
    E1 = tensor(1, **kw) if T1 is None else torch.exp(-dt / T1.to(dev))
    E2 = tensor(1, **kw) if T2 is None else torch.exp(-dt / T2.to(dev))
    Beff, γ, dt = (x.to(dev) for x in (Beff, γ, dt))
    E1, E2, γ, dt = (utils._tail(x, nd) for x in (E1, E2, γ, dt))
    E1_1, E2, g = E1 - 1, E2[..., None], 2 * π * γ * dt
    for t in range(Beff.shape[-2]):
        u, ϕ = beffective.beff2uϕ(Beff[..., t, :], g)
        Mr = utils.uϕrot(u, ϕ, M) if torch.any(ϕ != 0) else M
        M = _relax(Mr, E1, E1_1, E2)
    return M


def blochsim_ab(M: Tensor, A: Tensor, B: Tensor) -> Tensor:
    r"""``A @ M + B`` per spin (slowsims.py:117-131)."""
    return (A @ M[..., None]).squeeze(dim=-1) + B


def freeprec(M: Tensor, dur: Tensor, *, T1: Optional[Tensor] = None, T2: Optional[Tensor] = None,
             Δf: Optional[Tensor] = None) -> Tensor:
    r"""Free precession by plain autograd (slowsims.py:134-174); same arguments as ``sims.freeprec``."""
    nd = M.ndim
    dur = utils._tail(dur, nd)
    x, y, z = M.split(1, dim=-1)
    if Δf is not None:
        ϕ = -(2 * π) * utils._tail(Δf, nd) * dur
        c, s = torch.cos(ϕ), torch.sin(ϕ)
        x, y = c * x - s * y, s * x + c * y
    assert ((T1 is None) == (T2 is None))
    if T1 is not None:
        E1, E2 = torch.exp(-dur / utils._tail(T1, nd)), torch.exp(-dur / utils._tail(T2, nd))
        x, y, z = E2 * x, E2 * y, E1 * z + 1 - E1
    return torch.cat((x, y, z), dim=-1)
