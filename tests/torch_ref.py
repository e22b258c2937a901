"""Differentiable plain-torch statements of the stand-alone operators, used ONLY by the tests as the checker for the
CUDA kernels' forward values and explicit adjoints (autograd through these expressions gives every gradient).
Each follows the cited reference lines; nothing in the package imports this module."""
import math

import torch
import torch.nn.functional as F


def _tail(x, n):
    """Append singleton dims until x.ndim == n (no-op for 0-dim tensors)."""
    return x if x.ndim == 0 else x.reshape(x.shape + (1,) * (n - x.ndim))


def rfgr2beff(rf, gr, loc, Δf=None, b1Map=None, γ=None):
    """beffective.py:137-167 on (N,*Nd) grids: Bz = loc·gr + Δf/γ ; Bx + iBy = Σ_coils b1·rf."""
    N, Nd = loc.shape[0], tuple(loc.shape[1:-1])
    nd = len(Nd)
    Bz = torch.matmul(loc.reshape(N, -1, 3), gr).reshape((N,) + Nd + (-1,))
    if Δf is not None:
        Bz = Bz + _tail(Δf, nd + 2) / _tail(γ, nd + 2)
    rfx = rf.reshape((N,) + nd * (1,) + tuple(rf.shape[1:]))
    if b1Map is None:
        if rfx.ndim == Bz.ndim + 2:
            rfx = rfx.sum(dim=-1)
        Bx, By = rfx[..., 0, :].expand_as(Bz), rfx[..., 1, :].expand_as(Bz)
    else:
        b1 = b1Map if b1Map.ndim == nd + 3 else b1Map[..., None]
        if rfx.ndim == b1.ndim:
            rfx = rfx[..., None]
        br, bi = b1[..., 0, None, :], b1[..., 1, None, :]
        rx, ry = rfx[..., 0, :, :], rfx[..., 1, :, :]
        Bx = (br * rx - bi * ry).sum(dim=-1).expand_as(Bz)
        By = (br * ry + bi * rx).sum(dim=-1).expand_as(Bz)
    return torch.stack((Bx, By, Bz), dim=-1)


def beff2uphi(beff, g, dim=-1):
    """beffective.py:35-36."""
    return F.normalize(beff, dim=dim), -torch.norm(beff, dim=dim) * g


def _rodrigues(u, phi, v):
    """Rotate the vectors v (..., 3, k) about u (..., 3) by phi (...): utils.py:27-51."""
    c, s = torch.cos(phi)[..., None, None], torch.sin(phi)[..., None, None]
    u = u[..., None]
    return c * v + (1 - c) * u * (u * v).sum(dim=-2, keepdim=True) + s * torch.cross(u.expand_as(v), v, dim=-2)


def beff2ab(beff, E1, E2, γ, dt):
    """beffective.py:88-103: [A|B] <- relax(rotate([A|B])) for every step, B also recovering by 1-E1."""
    nd = beff.ndim - 2
    E1, E2, γ, dt = (_tail(x, nd) for x in (E1, E2, γ, dt))
    g = 2 * math.pi * γ * dt
    lead = beff.shape[:-2]
    AB = torch.eye(3, 4, dtype=beff.dtype, device=beff.device).expand(lead + (3, 4))
    one, zero = torch.ones_like(E1 * g), torch.zeros_like(E1 * g)
    scale = torch.stack((E2 * one, E2 * one, E1 * one), dim=-1)[..., None]
    rec = torch.stack((zero, zero, (1 - E1) * one), dim=-1)[..., None]
    pick_b = torch.tensor([0., 0., 0., 1.], dtype=beff.dtype, device=beff.device)
    for t in range(beff.shape[-2]):
        u, phi = beff2uphi(beff[..., t, :], g)
        AB = _rodrigues(u, phi, AB) * scale + rec * pick_b
    return AB[..., :3], AB[..., 3]


def freeprec(M, dur, T1=None, T2=None, Δf=None):
    """sims.py:345-369."""
    n = M.ndim - 1
    x, y, z = M.unbind(-1)
    if Δf is not None:
        ang = -2 * math.pi * _tail(Δf, n) * _tail(dur, n)
        c, s = torch.cos(ang), torch.sin(ang)
        x, y = c * x - s * y, s * x + c * y
    if T1 is not None:
        E1, E2 = torch.exp(-_tail(dur, n) / _tail(T1, n)), torch.exp(-_tail(dur, n) / _tail(T2, n))
        x, y, z = E2 * x, E2 * y, E1 * z + (1 - E1)
    return torch.stack((x, y, z), dim=-1)
