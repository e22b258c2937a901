r"""MRphy utilities: indexing helpers, k-space/gradient/slew conversions, RF re-parametrisations,
axis-angle rotation.  Same names and behaviour as ``/root/reference/mrphy/utils.py``; these are
O(nT) waveform-sized expressions.  The design-loop chain the optimiser differentiates through -- ``tρθ2rf`` /
``lρθ2rf``, ``ts2s``, ``s2g`` -- runs as ONE CUDA launch per call (and one for its adjoint) when its
tensors live on the GPU (``csrc/design_ops.cu``); ``ts2g`` and ``tρθts2rfgr`` fuse the whole chain into a single
launch.  CPU tensors, and the rare cases the kernel does not cover (gradients w.r.t. ``rfmax``/``smax``/``dt``, dtype
promotion, irregular broadcast shapes), evaluate the reference's torch expression, which is what these host utilities
are upstream.
"""
import os
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

_FLOATS = (torch.float32, torch.float64)


def _frozen(*ts) -> bool:
    return not any(t is not None and t.requires_grad for t in ts)


def _aux(x: Tensor, like: Tensor, rows: int, cols: int):
    """A small constant (rfmax / smax) as a `(rows|1, cols|1)` tensor on ``like``'s device and dtype, or None when its
    shape is not one the kernel broadcasts."""
    from mrphy import _ops
    v = _ops.on_device(x, like.device, like.dtype)
    if v.ndim == 0:
        v = v.reshape(1, 1)
    elif v.ndim == 1:
        v = v.reshape(-1, 1)                       # per pulse, like upstream's rfmax[:, None, None]
    if v.ndim != 2 or v.shape[0] not in (1, rows) or v.shape[1] not in (1, cols):
        return None
    return v


def _design_call(hr, hg, rf_kind: int, gr_kind: int):
    """One launch of the chain on prepared halves (either may be None) -> (rf, gr); the outputs remember their design
    variables so that `applypulse` can fold the chain's adjoint into its own gradient epilogue (`_ops.tag_design`)."""
    from mrphy import _ops
    rho, theta, rfmax = hr if hr is not None else (None, None, None)
    ts, smax, dt = hg if hg is not None else (None, None, None)
    rf, gr = _ops.design_waveform_cuda(rho, theta, rfmax, ts, smax, dt, rf_kind, gr_kind)
    _ops.tag_design(rf if rf_kind else None, gr if gr_kind else None, rho, theta, rfmax, ts, smax, dt, rf_kind, gr_kind)
    return rf, gr


def _rf_half(ρ: Tensor, θ: Tensor, rfmax: Tensor):
    """(ρ, θ, rfmax2) ready for the kernel, or None -> torch expression."""
    if os.environ.get('MRPHY_B200_REPARAM') == 'torch':      # measurement switch (profiles/design_step.py)
        return None
    if not (ρ.is_cuda and θ.is_cuda and ρ.dtype in _FLOATS and θ.dtype == ρ.dtype and ρ.shape == θ.shape
            and ρ.ndim in (3, 4) and ρ.shape[1] == 1 and ρ.numel() > 0 and _frozen(rfmax)):
        return None
    if rfmax.ndim > 0 and torch.result_type(ρ, rfmax) != ρ.dtype:
        return None
    if (rfmax.ndim == 2) != (ρ.ndim == 4) and rfmax.ndim > 0:   # upstream broadcasting needs (N,) with 3-D, (N,nCoils) with 4-D
        return None
    r = _aux(rfmax, ρ, ρ.shape[0], ρ.shape[3] if ρ.ndim == 4 else 1)
    return None if r is None else (ρ.contiguous(), θ.contiguous(), r)


def _gr_half(ts: Tensor, smax, dt):
    """(ts, smax2, dt) ready for the kernel (smax / dt may be None when unused), or None -> torch expression."""
    from mrphy import _ops
    if os.environ.get('MRPHY_B200_REPARAM') == 'torch':
        return None
    if not (ts.is_cuda and ts.dtype in _FLOATS and ts.ndim == 3 and ts.shape[1] == 3 and ts.numel() > 0
            and _frozen(smax, dt)):
        return None
    s2 = None
    if smax is not None:
        if smax.ndim > 0 and torch.result_type(ts, smax) != ts.dtype:
            return None
        if smax.ndim == 1:                       # smax[..., None] of a 1-D smax is per-axis (3,1) -> (1,3) here
            if smax.shape[0] not in (1, 3):
                return None
            s2 = _ops.on_device(smax, ts.device, ts.dtype).reshape(1, -1)
        else:
            s2 = _aux(smax, ts, ts.shape[0], 3)
        if s2 is None:
            return None
    d = None
    if dt is not None:
        if dt.numel() not in (1, ts.shape[0]) or (dt.ndim > 0 and torch.result_type(ts, dt) != ts.dtype):
            return None
        if dt.ndim > 1 and dt.shape[1:].numel() != 1:
            return None
        d = _ops.on_device(dt, ts.device)
    return ts.contiguous(), s2, d


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
    h = _gr_half(s, None, dt)
    if h is not None:
        return _design_call(None, h, 0, 2)[1]
    return _tail(dt, s.ndim) * torch.cumsum(s, dim=2)


def _unit(θ: Tensor) -> Tensor:
    return torch.cat((θ.cos(), θ.sin()), dim=1)


def lρθ2rf(lρ: Tensor, θ: Tensor, rfmax: Tensor) -> Tensor:
    r"""logit(ρ/rfmax), θ `(N,1,nT,(nCoils))` -> rf `(N,xy,nT,(nCoils))`."""
    h = _rf_half(lρ, θ, rfmax)
    if h is not None:
        return _design_call(h, None, 2, 0)[0]
    return lρ.sigmoid() * _per_pulse(rfmax) * _unit(θ)


def tρθ2rf(tρ: Tensor, θ: Tensor, rfmax: Tensor) -> Tensor:
    r"""tan(ρ/rfmax·π/2), θ `(N,1,nT,(nCoils))` -> rf `(N,xy,nT,(nCoils))`."""
    h = _rf_half(tρ, θ, rfmax)
    if h is not None:
        return _design_call(h, None, 1, 0)[0]
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
    r"""Scale samples with \|rf\| above ``rfmax-eps`` back onto that radius (one CUDA launch, and one for its adjoint, on
    GPU tensors)."""
    # (a limit of another dtype keeps the torch expression: upstream then forms rfmax - eps in THAT dtype)
    if rf.is_cuda and rf.dtype in _FLOATS and rf.ndim in (3, 4) and rf.shape[1] == 2 and rf.numel() > 0 and _frozen(rfmax) \
            and os.environ.get('MRPHY_B200_REPARAM') != 'torch' and rfmax.dtype == rf.dtype \
            and (rfmax.ndim == 0 or (rfmax.ndim == 2) == (rf.ndim == 4)):
        r = _aux(rfmax, rf, rf.shape[0], rf.shape[3] if rf.ndim == 4 else 1)
        if r is not None:
            from mrphy import _ops
            return _ops.clamp_waveform_cuda(rf.contiguous(), r, float(eps), 1)
    scale = ((_per_pulse(rfmax) - eps) / rf.norm(dim=1, keepdim=True)).clamp_(max=1)
    return rf.mul(scale)


def s2ts(s: Tensor, smax: Tensor) -> Tensor:
    r"""slew `(N,xyz,nT)` -> tan(s/smax·π/2)."""
    return (s / smax[..., None] * π / 2).tan()


def ts2s(ts: Tensor, smax: Tensor) -> Tensor:
    r"""tan(s/smax·π/2) -> slew `(N,xyz,nT)`."""
    h = _gr_half(ts, smax, None)
    if h is not None:
        return _design_call(None, h, 0, 3)[1]
    return ts.atan() / π * 2 * smax[..., None]


def ts2g(ts: Tensor, smax: Tensor, dt: Tensor = dt0) -> Tensor:
    r"""``s2g(ts2s(ts, smax), dt)`` in one launch (not in the reference; the chain every slew-constrained design runs)."""
    h = _gr_half(ts, smax, dt)
    if h is not None:
        return _design_call(None, h, 0, 1)[1]
    return s2g(ts2s(ts, smax), dt)


def tρθts2rfgr(tρ: Tensor, θ: Tensor, ts: Tensor, rfmax: Tensor, smax: Tensor, dt: Tensor = dt0, *,
               logit: bool = False) -> Tuple[Tensor, Tensor]:
    r"""The whole re-parametrisation of a design step, ``(tρθ2rf | lρθ2rf)(ρ, θ, rfmax)`` and
    ``s2g(ts2s(ts, smax), dt)``, as ONE launch forward and ONE backward -> ``(rf, gr)``."""
    hr, hg = _rf_half(tρ, θ, rfmax), _gr_half(ts, smax, dt)
    if hr is not None and hg is not None and hr[0].dtype == hg[0].dtype and hr[0].shape[0] == hg[0].shape[0] \
            and hr[0].shape[2] == hg[0].shape[2] and hr[0].device == hg[0].device:
        return _design_call(hr, hg, 2 if logit else 1, 1)
    return (lρθ2rf if logit else tρθ2rf)(tρ, θ, rfmax), ts2g(ts, smax, dt)


def sclamp(s: Tensor, smax: Tensor) -> Tensor:
    r"""Clamp slew rate componentwise to ``±smax`` `(N,xyz)` (one CUDA launch, and one for its adjoint, on GPU tensors)."""
    if s.is_cuda and s.dtype in _FLOATS and s.ndim == 3 and s.shape[1] == 3 and s.numel() > 0 and _frozen(smax) \
            and os.environ.get('MRPHY_B200_REPARAM') != 'torch' and smax.ndim <= 2:
        m = _aux(smax if smax.ndim != 1 else smax.reshape(1, -1), s, s.shape[0], 3)
        if m is not None:
            from mrphy import _ops
            return _ops.clamp_waveform_cuda(s.contiguous(), m, 0.0, 2)
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
