r"""MRphy utilities: indexing helpers, k-space/gradient/slew conversions, RF re-parametrisations,
axis-angle rotation.  Same names and behaviour as ``/root/reference/mrphy/utils.py``; these are
O(nT) waveform-sized torch expressions (not part of the CUDA hot path) and run on any device.
"""
from numbers import Number
from typing import Any, Tuple, Union

import numpy as np
import torch
from numpy import ndarray as ndarray_c
from torch import Tensor

from mrphy import γH, dt0, π, __CUPY_IS_AVAILABLE__
if __CUPY_IS_AVAILABLE__:
    import cupy as cp
    from cupy import ndarray as ndarray_g
    ndarrayA = Union[ndarray_c, ndarray_g]
else:
    ndarrayA = ndarray_c

__all__ = ['ctrsub', 'g2k', 'g2s', 'k2g', 'rf_c2r', 'rf_r2c', 'rf2tρθ',
           'rfclamp', 's2g', 's2ts', 'sclamp', 'ts2s', 'tρθ2rf', 'uφrot']


def _tail(x: Tensor, ndim: int) -> Tensor:
    """Right-pad the shape of ``x`` with singleton dims up to ``ndim`` dims."""
    return x.reshape(x.shape + (ndim - x.ndim) * (1,))


def _first_diff(x: Tensor) -> Tensor:
    """x[0], x[1]-x[0], x[2]-x[1], ... along dim 2."""
    return torch.cat((x[:, :, :1], x[:, :, 1:] - x[:, :, :-1]), dim=2)


def _per_pulse(rfmax: Tensor) -> Tensor:
    """rfmax () or (N,(nCoils)) -> (N,1,1,(nCoils))."""
    rfmax = rfmax[None] if rfmax.ndim == 0 else rfmax
    return rfmax[:, None, None, ...]


def ctrsub(shape: Any) -> Any:
    r"""Center subscript of a regular grid: ``shape//2`` (utils.py:27-33)."""
    return shape // 2


def g2k(g: Tensor, isTx: bool, dt: Tensor = dt0, *, γ: Tensor = γH) -> Tensor:
    r"""Gradient `(N,xyz,nT)` G/cm -> k-space cycle/cm; a transmit k-space ends at the origin."""
    γ, dt = _tail(γ, g.ndim), _tail(dt, g.ndim)
    k = γ * dt * torch.cumsum(g, dim=2)
    return k - k[:, :, [-1]] if isTx else k


def g2s(g: Tensor, dt: Tensor = dt0) -> Tensor:
    r"""Gradient `(N,xyz,nT)` -> slew rate (first difference / dt)."""
    return _first_diff(g) / _tail(dt, g.ndim)


def k2g(k: Tensor, isTx: bool, dt: Tensor = dt0, *, γ: Tensor = γH) -> Tensor:
    r"""k-space `(N,xyz,nT)` -> gradient; a transmit k-space must end at 0."""
    assert ((not isTx) or torch.all(k[:, :, -1] == 0))
    γ, dt = _tail(γ, k.ndim), _tail(dt, k.ndim)
    return _first_diff(k) / γ / dt


def s2g(s: Tensor, dt: Tensor = dt0) -> Tensor:
    r"""Slew rate `(N,xyz,nT)` -> gradient (running sum * dt)."""
    return _tail(dt, s.ndim) * torch.cumsum(s, dim=2)


def _unit(θ: Tensor) -> Tensor:
    return torch.cat((θ.cos(), θ.sin()), dim=1)


def lρθ2rf(lρ: Tensor, θ: Tensor, rfmax: Tensor) -> Tensor:
    r"""logit(ρ/rfmax), θ `(N,1,nT,(nCoils))` -> rf `(N,xy,nT,(nCoils))`."""
    return lρ.sigmoid() * _per_pulse(rfmax) * _unit(θ)


def tρθ2rf(tρ: Tensor, θ: Tensor, rfmax: Tensor) -> Tensor:
    r"""tan(ρ/rfmax·π/2), θ `(N,1,nT,(nCoils))` -> rf `(N,xy,nT,(nCoils))`."""
    return tρ.atan() / π * 2 * _per_pulse(rfmax) * _unit(θ)


def _phase(rf: Tensor) -> Tensor:
    return torch.atan2(rf[:, [1], :], rf[:, [0], :])


def rf2lρθ(rf: Tensor, rfmax: Tensor, *, eps: Number = 1e-7) -> Tuple[Tensor, Tensor]:
    r"""rf `(N,xy,nT,(nCoils))` -> (logit(|rf|/rfmax), phase)."""
    return (rf.norm(dim=1, keepdim=True) / _per_pulse(rfmax)).logit(eps), _phase(rf)


def rf2tρθ(rf: Tensor, rfmax: Tensor) -> Tuple[Tensor, Tensor]:
    r"""rf `(N,xy,nT,(nCoils))` -> (tan(|rf|/rfmax·π/2), phase)."""
    return (rf.norm(dim=1, keepdim=True) / _per_pulse(rfmax) * π / 2).tan(), _phase(rf)


def rfclamp(rf: Tensor, rfmax: Tensor, *, eps: Number = 1e-7) -> Tensor:
    r"""Scale samples with \|rf\| above ``rfmax-eps`` back onto that radius."""
    scale = ((_per_pulse(rfmax) - eps) / rf.norm(dim=1, keepdim=True)).clamp_(max=1)
    return rf.mul(scale)


def s2ts(s: Tensor, smax: Tensor) -> Tensor:
    r"""slew `(N,xyz,nT)` -> tan(s/smax·π/2)."""
    return (s / smax[..., None] * π / 2).tan()


def ts2s(ts: Tensor, smax: Tensor) -> Tensor:
    r"""tan(s/smax·π/2) -> slew `(N,xyz,nT)`."""
    return ts.atan() / π * 2 * smax[..., None]


def sclamp(s: Tensor, smax: Tensor) -> Tensor:
    r"""Clamp slew rate componentwise to ``±smax`` `(N,xyz)`."""
    smax = (smax[None] if smax.ndim == 0 else smax).to(s)[..., None]
    return s.max(-smax).min(smax)


def rf_c2r(rf: ndarrayA) -> ndarrayA:
    r"""complex rf `(N,1,nT,(nCoils))` -> real `(N,xy,nT,(nCoils))` (numpy or cupy)."""
    if isinstance(rf, ndarray_c):
        return np.concatenate((np.real(rf), np.imag(rf)), axis=1)
    if __CUPY_IS_AVAILABLE__:
        return cp.concatenate((cp.real(rf), cp.imag(rf)), axis=1)
    raise TypeError(f'Unknown type: {type(rf)}')


def rf_r2c(rf: ndarrayA) -> ndarrayA:
    r"""real rf `(N,xy,nT,(nCoils))` -> complex `(N,1,nT,(nCoils))`."""
    return rf[:, [0], ...] + 1j * rf[:, [1], ...]


def uϕrot(U: Tensor, Φ: Tensor, Vi: Tensor) -> Tensor:
    r"""Rotate ``Vi`` `(N,*Nd,xyz,(nV))` about unit axes ``U`` `(N,*Nd,xyz)` by ``Φ`` `(N,*Nd)` (Rodrigues)."""
    if Vi.ndim == U.ndim:
        dim, Φ = -1, Φ[..., None]
    else:
        dim, Φ, U = -2, Φ[..., None, None], U[..., None]
    c, s = torch.cos(Φ), torch.sin(Φ)
    along = torch.sum(U * Vi, dim=dim, keepdim=True) * U
    return c * Vi + (1 - c) * along + s * torch.cross(U.expand_as(Vi), Vi, dim=dim)
