r"""Objects for MRI excitation simulation: ``Pulse``, ``SpinArray``, ``SpinCube``, ``Examples``.

Same public behaviour as ``/root/reference/mrphy/mobjs.py`` (constructor keywords incl. the Greek
ones, attribute coercion rules, ``asdict`` keys, compact ``_`` attributes with mask embed/extract,
``interpT`` grid), re-implemented.  ``SpinArray.applypulse`` -- the entry of the hot path -- calls the
fused CUDA op (``_ops.fused_applypulse``) instead of materialising ``Beff`` and looping over time in
Python (mobjs.py:435-446 upstream).
"""
import copy
from typing import Optional

import numpy as np
import torch
from torch import tensor, Tensor

from mrphy import γH, dt0, gmax0, smax0, rfmax0, T1G, T2G, π
from mrphy import utils, beffective, sims, _ops

__all__ = ['Pulse', 'SpinArray', 'SpinCube', 'Examples']


_CONSTS = {id(c): c for c in (γH, dt0, gmax0, smax0, rfmax0, T1G, T2G)}
_const_cache = {}


def _as_tensor(v, device, dtype) -> Tensor:
    if not isinstance(v, Tensor):
        return tensor(v, device=device, dtype=dtype)
    if id(v) in _CONSTS and device.type == 'cuda':
        # package defaults: one (synchronising) host->device copy per device/dtype; objects get device-side clones
        key = (id(v), device, dtype)
        hit = _const_cache.get(key)
        if hit is None:
            hit = _const_cache[key] = v.to(device=device, dtype=dtype)
        return hit.clone()
    return v.to(device=device, dtype=dtype)


class _Slotted(object):
    """Objects whose attribute writes are validated; ``copy.deepcopy`` bypasses the validation."""
    __slots__ = ()

    def _all_slots(self):
        names = []
        for klass in type(self).__mro__:
            names.extend(getattr(klass, '__slots__', ()))
        return names

    def __deepcopy__(self, memo):
        new = object.__new__(type(self))
        memo[id(self)] = new
        for k in self._all_slots():
            try:
                v = object.__getattribute__(self, k)
            except AttributeError:
                continue
            object.__setattr__(new, k, copy.deepcopy(v, memo))
        return new


class Pulse(_Slotted):
    r"""Pulse object of RF and GR.

    Usage:
        ``pulse = Pulse(rf, gr, *, dt, gmax, smax, rfmax, desc, device, dtype)``

    Inputs (at least one of ``rf``, ``gr``; the other defaults to zeros):
        - ``rf``: `(N,xy,nT,(nCoils))` "Gauss";  ``gr``: `(N,xyz,nT)` "Gauss/cm"
        - ``dt``: `()` ⊻ `(N ⊻ 1,)` "Sec";  ``gmax``, ``smax``: `()` ⊻ `(N ⊻ 1, xyz ⊻ 1)`;
          ``rfmax``: `()` ⊻ `(N ⊻ 1,(nCoils))`;  ``desc``: str;  ``device``; ``dtype``

    Properties: ``device``, ``dtype``, ``is_cuda``, ``shape`` = `(N,1,nT)` (read-only);
    ``rf``, ``gr``, ``dt`` `(N ⊻ 1,)`, ``gmax``/``smax`` `(N ⊻ 1, xyz)`, ``rfmax`` `(N ⊻ 1,(nCoils))`, ``desc``.
    Every tensor assigned is moved to the pulse's device and dtype.
    """

    _readonly = ('device', 'dtype', 'is_cuda', 'shape')
    _limits = ('gmax', 'smax', 'rfmax')
    __slots__ = set(_readonly + _limits + ('rf', 'gr', 'dt', 'desc'))

    def __init__(
        self,
        rf: Optional[Tensor] = None, gr: Optional[Tensor] = None, *,
        dt: Tensor = dt0,
        gmax: Tensor = gmax0, smax: Tensor = smax0, rfmax: Tensor = rfmax0,
        desc: str = "generic pulse",
        device: torch.device = torch.device('cpu'),
        dtype: torch.dtype = torch.float32
    ):
        assert (isinstance(device, torch.device) and isinstance(dtype, torch.dtype))
        assert not (rf is None and gr is None), "Missing both `rf` and `gr` inputs"
        for k, v in (('device', device), ('dtype', dtype), ('is_cuda', device.type == 'cuda')):
            object.__setattr__(self, k, v)
        kw = {'device': device, 'dtype': dtype}
        ref = rf if rf is not None else gr
        N, nT = ref.shape[0], ref.shape[2]
        if rf is None:
            rf = torch.zeros((N, 2, nT), **kw)
        if gr is None:
            gr = torch.zeros((N, 3, nT), **kw)
        assert (N == gr.shape[0] and nT == gr.shape[2])
        object.__setattr__(self, 'shape', torch.Size((N, 1, nT)))
        self.rf, self.gr = rf.to(**kw), gr.to(**kw)
        self.dt, self.gmax, self.smax, self.rfmax = dt, gmax, smax, rfmax
        self.desc = desc

    def __setattr__(self, k, v):
        if k in self._readonly:
            raise AttributeError(f"'Pulse' object attribute '{k}' is read-only")
        if k != 'desc':
            v = _as_tensor(v, self.device, self.dtype)
        if k in ('rf', 'gr'):
            assert (v.shape[0] == self.shape[0] and v.shape[2] == self.shape[2])
        elif k in ('gmax', 'smax'):       # -> (N ⊻ 1, xyz)
            v = v.expand((1 if v.ndim == 0 else v.shape[0], self.gr.shape[1]))
        elif k == 'rfmax':                # -> (N ⊻ 1,(nCoils))
            if v.ndim == 0:
                v = v[None]
            elif v.ndim == 2 and v.shape[1] == 1:
                v = v[:, 0]
        elif k == 'dt':                   # -> (N ⊻ 1,)
            if v.ndim == 0:
                v = v[None]
            assert (v.ndim == 1)
        object.__setattr__(self, k, v)

    def asdict(self, *, toNumpy: bool = True) -> dict:
        r"""``d = pulse.asdict(*, toNumpy)``: detached copies of the data; ``Pulse(**d)`` rebuilds it."""
        conv = (lambda x: x.detach().cpu().numpy()) if toNumpy else (lambda x: x.detach())
        d = {k: conv(getattr(self, k)) for k in ('rf', 'gr', 'dt', 'gmax', 'smax', 'rfmax')}
        d.update({k: getattr(self, k) for k in ('desc', 'device', 'dtype')})
        return d

    def beff(self, loc: Tensor, *, Δf: Optional[Tensor] = None, b1Map: Optional[Tensor] = None,
             γ: Tensor = γH) -> Tensor:
        r"""``beff = pulse.beff(loc, *, Δf, b1Map, γ)``: B-effective `(N,*Nd,nT,xyz)` at ``loc`` `(N,*Nd,xyz)`."""
        mv = lambda x: None if x is None else x.to(device=self.device)
        return beffective.rfgr2beff(self.rf, self.gr, mv(loc), Δf=mv(Δf), b1Map=mv(b1Map), γ=mv(γ))

    def interpT(self, dt: Tensor, *, kind: str = 'linear') -> 'Pulse':
        r"""``new_pulse = pulse.interpT(dt, *, kind)``: resample to dwell time ``dt`` `(1,)`.

        Samples are end-of-interval (sample k sits at (k+1)·dt); a zero sample is assumed at t=0.  The
        new length is ``t_end // dt_new`` evaluated in Python floats exactly as upstream (mobjs.py:211-212),
        so e.g. nT=10, 4µs→2µs yields 19 samples.  Non-differentiable; limits are reset to defaults.
        """
        assert (self.dt.numel() == dt.numel() == 1)
        dt_o, dt_n = self.dt.item(), dt.item()
        if dt_o == dt_n:
            return copy.deepcopy(self)
        kw = {'device': self.device, 'dtype': self.dtype}
        nT = self.shape[2]
        t_o = np.arange(0, nT + 1) * dt_o
        t_n = np.arange(1, t_o[-1] // dt_n + 1) * dt_n

        def resample(x: Tensor) -> Tensor:
            x0 = torch.cat((torch.zeros_like(x[:, :, :1]), x.detach()), dim=2)
            if kind == 'linear':   # on device: gather the two bracketing samples
                hi = np.clip(np.searchsorted(t_o, t_n, side='left'), 1, nT)
                w = (t_n - t_o[hi - 1]) / (t_o[hi] - t_o[hi - 1])
                shp = (1, 1, -1) + (1,) * (x.ndim - 3)
                w = torch.as_tensor(w, device=x.device, dtype=torch.float64).reshape(shp)
                hi = torch.as_tensor(hi, device=x.device)
                lo_v, hi_v = x0.index_select(2, hi - 1).double(), x0.index_select(2, hi).double()
                return (lo_v + (hi_v - lo_v) * w).to(**kw)
            from scipy import interpolate   # other kinds: host round trip like upstream
            f = interpolate.interp1d(t_o, x0.cpu().numpy(), axis=2, kind=kind, copy=False, assume_sorted=True)
            return tensor(f(t_n), **kw)

        desc = f"{self.desc} + interpT\'ed: dt = {dt_n}"
        return Pulse(resample(self.rf), resample(self.gr), dt=dt, desc=desc, **kw)

    def to(self, *, device: torch.device = torch.device('cpu'), dtype: torch.dtype = torch.float32) -> 'Pulse':
        r"""The same pulse on ``device`` with ``dtype`` (``self`` when nothing changes)."""
        if self.device == device and self.dtype == dtype:
            return self
        return Pulse(self.rf, self.gr, dt=self.dt, desc=self.desc, device=device, dtype=dtype)


class SpinArray(_Slotted):
    r"""mrphy.mobjs.SpinArray object.

    Usage:
        ``spinarray = SpinArray(shape, mask, *, T1_ ⊻ T1, T2_ ⊻ T2, γ_ ⊻ γ, M_ ⊻ M, device, dtype)``

    ``shape`` = `(N, *Nd)`; ``mask`` `(1, *Nd)` bool selects the ``nM`` spins kept in the compact
    ``_`` attributes: ``T1_``, ``T2_``, ``γ_`` `(N, nM)` and ``M_`` `(N, nM, xyz)`.  The non-compact
    names (``T1``, ``T2``, ``γ``, ``M``) embed on read and extract on write.  The mask is global to
    the batch and must not be modified; indexed assignment into a non-compact attribute does not
    write through (use :meth:`crds_`).
    """

    _readonly = ('shape', 'mask', 'device', 'dtype', 'is_cuda', 'ndim', 'nM')
    _compact = ('T1_', 'T2_', 'γ_', 'M_')
    __slots__ = set(_readonly + _compact)

    def __init__(
        self, shape: tuple, mask: Optional[Tensor] = None, *,
        T1: Optional[Tensor] = None, T1_: Optional[Tensor] = None,
        T2: Optional[Tensor] = None, T2_: Optional[Tensor] = None,
        γ: Optional[Tensor] = None, γ_: Optional[Tensor] = None,
        M: Optional[Tensor] = None, M_: Optional[Tensor] = None,
        device: torch.device = torch.device('cpu'),
        dtype: torch.dtype = torch.float32
    ):
        shape = tuple(shape)
        mask = (torch.ones((1,) + shape[1:], dtype=torch.bool, device=device)
                if mask is None else mask.to(device=device))
        assert (isinstance(device, torch.device) and isinstance(dtype, torch.dtype) and
                mask.dtype == torch.bool and mask.shape == (1,) + shape[1:])
        for k, v in (('shape', shape), ('mask', mask), ('ndim', len(shape)),
                     ('nM', torch.count_nonzero(mask).item()), ('device', device), ('dtype', dtype),
                     ('is_cuda', device.type == 'cuda')):
            object.__setattr__(self, k, v)
        defaults = {'T1': T1G, 'T2': T2G, 'γ': γH, 'M': tensor([0., 0., 1.])}
        given = {'T1': (T1, T1_), 'T2': (T2, T2_), 'γ': (γ, γ_), 'M': (M, M_)}
        for k, (full, compact) in given.items():
            assert ((full is None) or (compact is None))
            if full is not None:
                setattr(self, k, full)
            else:
                setattr(self, k + '_', defaults[k] if compact is None else compact)

    def _is_full(self) -> bool:
        return self.nM == int(np.prod(self.shape[1:]))

    def __getattr__(self, k):   # only reached when normal lookup fails: the non-compact views
        if k + '_' not in self._compact:
            raise AttributeError(f"'SpinArray' has no attribute '{k}'")
        v_ = getattr(self, k + '_')
        return v_.reshape(self.shape + v_.shape[2:]) if self._is_full() else self.embed(v_)

    def __setattr__(self, k_, v_):
        if k_ in self._readonly:
            raise AttributeError(f"'SpinArray' object attribute '{k_}' is read-only")
        v_ = _as_tensor(v_, self.device, self.dtype)
        shape = self.shape
        if k_ + '_' in self._compact:    # non-compact assignment
            k_ = k_ + '_'
            v_ = self.extract(v_.expand(shape + (3,) if k_ == 'M_' else shape))
        if k_ == 'M_':
            want = shape[:1] + (self.nM, 3)
            if v_.shape != want:
                v_ = v_.expand(want).clone()
        elif k_ in self._compact:
            v_ = v_.expand((shape[0], self.nM))
        object.__setattr__(self, k_, v_)

    def applypulse(
        self, pulse: Pulse, *,
        doEmbed: bool = False, doRelax: bool = True, doUpdate: bool = False,
        loc: Optional[Tensor] = None, loc_: Optional[Tensor] = None,
        Δf: Optional[Tensor] = None, Δf_: Optional[Tensor] = None,
        b1Map: Optional[Tensor] = None, b1Map_: Optional[Tensor] = None
    ) -> Tensor:
        r"""Apply a pulse to the spinarray object.

        Usage:
            ``M = spinarray.applypulse(pulse, *, loc, doEmbed=True, doRelax, doUpdate, Δf, b1Map)``
            ``M_ = spinarray.applypulse(pulse, *, loc_, doEmbed=False, doRelax, doUpdate, Δf_, b1Map_)``
        Inputs:
            - ``pulse``: mrphy.mobjs.Pulse;  ``loc`` ⊻ ``loc_``: `(N,*Nd ⊻ nM,xyz)` "cm"
        Optionals:
            - ``doEmbed`` [t/F]: return ``M`` or ``M_``;  ``doRelax`` [T/f];  ``doUpdate`` [t/F]: store the
              result in ``self.M_`` (keeps the autograd graph, as upstream)
            - ``Δf`` ⊻ ``Δf_``: `(N,*Nd ⊻ nM)` "Hz";  ``b1Map`` ⊻ ``b1Map_``: `(N,*Nd ⊻ nM,xy,(nCoils))`
        Outputs:
            - ``M`` ⊻ ``M_``: `(N,*Nd ⊻ nM,xyz)`

        One fused CUDA kernel simulates all ``nT`` steps; gradients flow to ``pulse.rf``, ``pulse.gr`` and
        ``self.M_``.  When ``loc``/``Δf``/``b1Map`` themselves require grad the field is materialised with
        ``pulse2beff`` and the explicit-``Beff`` kernels are used, which reproduces upstream's autograd.
        """
        assert ((loc_ is None) != (loc is None))
        loc_ = loc_ if loc is None else self.extract(loc)
        assert ((Δf_ is None) or (Δf is None))
        Δf_ = Δf_ if Δf is None else self.extract(Δf)
        assert ((b1Map_ is None) or (b1Map is None))
        b1Map_ = b1Map_ if b1Map is None else self.extract(b1Map)
        T1_, T2_ = (self.T1_, self.T2_) if doRelax else (None, None)

        geom_grad = any(x is not None and x.requires_grad for x in (loc_, Δf_, b1Map_, self.γ_))
        if geom_grad and torch.is_grad_enabled():
            beff_ = self.pulse2beff(pulse, loc_=loc_, Δf_=Δf_, b1Map_=b1Map_, doEmbed=False)
            M_ = sims.blochsim(self.M_, beff_, T1=T1_, T2=T2_, γ=self.γ_, dt=pulse.dt)
        else:
            pulse = pulse.to(device=self.device, dtype=self.dtype)
            M_ = _ops.fused_applypulse(self.M_, pulse.rf, pulse.gr, loc_, Δf_=Δf_, b1Map_=b1Map_,
                                       T1_=T1_, T2_=T2_, γ_=self.γ_, dt=pulse.dt)
        if doUpdate:
            self.M_ = M_
        return self.embed(M_) if doEmbed else M_

    def asdict(self, *, toNumpy: bool = True, doEmbed: bool = True) -> dict:
        r"""``d = spinarray.asdict(*, toNumpy, doEmbed)``: detached data + ``mask, shape, device, dtype``."""
        conv = (lambda x: x.detach().cpu().numpy()) if toNumpy else (lambda x: x.detach())
        keys = ('T1', 'T2', 'γ', 'M') if doEmbed else ('T1_', 'T2_', 'γ_', 'M_')
        d = {k: conv(getattr(self, k)) for k in keys}
        d['mask'] = conv(self.mask)
        d.update({k: getattr(self, k) for k in ('shape', 'device', 'dtype')})
        return d

    def crds_(self, crds: list) -> list:
        r"""``crds_ = spinarray.crds_(crds)``: translate indices into `(N,*Nd,...)` arrays to indices into the
        compact `(N,nM,...)` arrays, so that ``v_[crds_] == v[crds]`` and ``v_[crds_] = x`` writes through."""
        mask, ndim, nM = self.mask, self.ndim, self.nM
        assert (len(crds) >= ndim)
        lut = torch.full(mask.shape, -1, dtype=torch.int64)
        lut[mask.cpu()] = torch.arange(nM)
        picked = lut[tuple([[0]] + list(crds[1:ndim]))].tolist()
        inds_ = [i for i in picked if i != -1]
        return [crds[0], inds_] + [crds[i] for i in range(ndim, len(crds))]

    def dim(self) -> int:
        r"""``len(spinarray.shape)``."""
        return len(self.shape)

    def embed(self, v_: Tensor, *, out: Optional[Tensor] = None) -> Tensor:
        r"""``out = spinarray.embed(v_, *, out)``: compact `(N,nM,...)` -> `(N,*Nd,...)`, NaN outside the mask."""
        out = v_.new_full(self.shape + v_.shape[2:], float('NaN')) if out is None else out
        out[self.mask.expand(self.shape)] = v_.reshape((-1,) + v_.shape[2:])
        return out

    def extract(self, v: Tensor, *, out_: Optional[Tensor] = None) -> Tensor:
        r"""``out_ = spinarray.extract(v, *, out_)``: `(N,*Nd,...)` -> compact `(N,nM,...)` (row-major order)."""
        picked = v[self.mask.expand(self.shape)]
        oshape = (self.shape[0], self.nM) + v.shape[self.ndim:]
        if out_ is None:
            return picked.reshape(oshape)
        out_.view((-1,) + v.shape[self.ndim:]).copy_(picked)
        return out_

    def freeprec(
        self, dur: Tensor, *,
        doEmbed: bool = False, doRelax: bool = True, doUpdate: bool = False,
        Δf: Optional[Tensor] = None, Δf_: Optional[Tensor] = None
    ) -> Tensor:
        r"""``M(_) = obj.freeprec(dur, *, doEmbed, doRelax, doUpdate, Δf ⊻ Δf_)``: free precession for ``dur`` s."""
        assert ((Δf_ is None) or (Δf is None))
        Δf_ = Δf_ if Δf is None else self.extract(Δf)
        T1_, T2_ = (self.T1_, self.T2_) if doRelax else (None, None)
        M_ = sims.freeprec(self.M_, dur.to(self.device), T1=T1_, T2=T2_, Δf=Δf_)
        if doUpdate:
            self.M_ = M_
        return self.embed(M_) if doEmbed else M_

    def mask_(self, *, mask: Tensor) -> Tensor:
        r"""``mask_ = spinarray.mask_(mask=mask)``: an external ``mask`` `(1,*Nd)` restricted to the compact
        spins, `(1,nM)`.  (Upstream's version calls a tensor and always raises, mobjs.py:605.)"""
        return mask.to(self.device)[self.mask].reshape((1, -1))

    def numel(self) -> int:
        r"""Number of grid points incl. masked-out ones (``mask.numel()``)."""
        return self.mask.numel()

    def pulse2beff(
        self, pulse: Pulse, *, doEmbed: bool = False,
        loc: Optional[Tensor] = None, loc_: Optional[Tensor] = None,
        Δf: Optional[Tensor] = None, Δf_: Optional[Tensor] = None,
        b1Map: Optional[Tensor] = None, b1Map_: Optional[Tensor] = None
    ) -> Tensor:
        r"""``beff(_) = spinarray.pulse2beff(pulse, *, loc ⊻ loc_, doEmbed, Δf ⊻ Δf_, b1Map ⊻ b1Map_)``:
        B-effective `(N,*Nd ⊻ nM,nT,xyz)` of ``pulse`` with this object's ``γ``."""
        assert ((loc_ is None) != (loc is None))
        loc_ = loc_ if loc is None else self.extract(loc)
        assert ((Δf_ is None) or (Δf is None))
        Δf_ = Δf_ if Δf is None else self.extract(Δf)
        assert ((b1Map_ is None) or (b1Map is None))
        b1Map_ = b1Map_ if b1Map is None else self.extract(b1Map)
        pulse = pulse.to(device=self.device, dtype=self.dtype)
        beff_ = pulse.beff(loc_, γ=self.γ_, Δf=Δf_, b1Map=b1Map_)
        return self.embed(beff_) if doEmbed else beff_

    def size(self) -> tuple:
        r"""``spinarray.shape``."""
        return self.shape

    def to(self, *, device: torch.device = torch.device('cpu'),
           dtype: torch.dtype = torch.float32) -> 'SpinArray':
        r"""The same spin array on ``device`` with ``dtype`` (``self`` when nothing changes)."""
        if self.device == device and self.dtype == dtype:
            return self
        return SpinArray(self.shape, self.mask, T1_=self.T1_, T2_=self.T2_, γ_=self.γ_, M_=self.M_,
                         device=device, dtype=dtype)


class SpinCube(SpinArray):
    r"""mrphy.mobjs.SpinCube object: a SpinArray on a regular grid.

    Usage:
        ``SpinCube(shape, fov, *, mask, ofst, Δf_ ⊻ Δf, T1_ ⊻ T1, T2_ ⊻ T2, γ_ ⊻ γ, M_ ⊻ M, device, dtype)``

    ``fov``, ``ofst``: `(N, xyz)` "cm".  Extra properties: ``spinarray`` (the SpinArray part),
    ``Δf_`` `(N, nM)` "Hz", ``loc_`` `(N, nM, xyz)` "cm" (read-only, recomputed whenever ``fov`` or ``ofst``
    is set: ``loc = fov·(idx − n//2)/n + ofst``).  Unknown attributes are forwarded to ``spinarray``.
    """

    _readonly = ('spinarray', 'loc_')
    _compact = ('Δf_', 'loc_')
    __slots__ = set(_readonly + _compact + ('fov', 'ofst'))

    def __init__(
        self, shape: tuple, fov: Tensor, *, mask: Optional[Tensor] = None,
        ofst: Tensor = tensor([[0., 0., 0.]]),
        Δf: Optional[Tensor] = None, Δf_: Optional[Tensor] = None,
        T1: Optional[Tensor] = None, T1_: Optional[Tensor] = None,
        T2: Optional[Tensor] = None, T2_: Optional[Tensor] = None,
        γ: Optional[Tensor] = None, γ_: Optional[Tensor] = None,
        M: Optional[Tensor] = None, M_: Optional[Tensor] = None,
        device: torch.device = torch.device('cpu'),
        dtype: torch.dtype = torch.float32
    ):
        sp = SpinArray(shape, mask, T1=T1, T1_=T1_, T2=T2, T2_=T2_, γ=γ, γ_=γ_, M=M, M_=M_,
                       device=device, dtype=dtype)
        object.__setattr__(self, 'spinarray', sp)
        kw = {'device': sp.device, 'dtype': sp.dtype}
        object.__setattr__(self, 'fov', fov.to(**kw))
        object.__setattr__(self, 'ofst', ofst.to(**kw))
        object.__setattr__(self, 'loc_', torch.zeros((sp.shape[0], sp.nM, 3), **kw))
        self._update_loc_()
        assert ((Δf is None) or (Δf_ is None))
        if Δf is None:
            self.Δf_ = tensor(0.) if Δf_ is None else Δf_
        else:
            self.Δf = Δf

    def __getattr__(self, k):
        if k + '_' not in self._compact:
            # object.__getattribute__ (not self.spinarray) so a half-built copy cannot recurse forever
            sp = object.__getattribute__(self, 'spinarray')
            try:
                return getattr(sp, k)
            except AttributeError:
                raise AttributeError(f"'SpinCube' has no attribute '{k}'")
        v_, sp = getattr(self, k + '_'), self.spinarray
        return v_.reshape(sp.shape + v_.shape[2:]) if sp._is_full() else sp.embed(v_)

    def __setattr__(self, k_, v_):
        if (k_ in self._readonly) or (k_ + '_' in self._readonly):
            raise AttributeError(f"'SpinCube' object attribute '{k_}' is read-only")
        sp = self.spinarray
        if k_ in SpinArray.__slots__ or k_ + '_' in SpinArray.__slots__:
            setattr(sp, k_, v_)
            return
        v_ = _as_tensor(v_, sp.device, sp.dtype)
        if k_ + '_' in self._compact:
            k_ = k_ + '_'
            v_ = sp.extract(v_.expand(sp.shape))
        if k_ == 'Δf_':
            v_ = v_.expand((sp.shape[0], sp.nM))
        elif k_ in ('fov', 'ofst'):
            assert (v_.ndim == 2)
        object.__setattr__(self, k_, v_)
        if k_ in ('fov', 'ofst'):
            self._update_loc_()

    def _update_loc_(self):
        r"""Recompute ``loc_`` in place from ``fov`` and ``ofst`` (grid centred on index ``n//2``)."""
        sp = self.spinarray
        kw = {'device': sp.device, 'dtype': sp.dtype}
        axes = [(torch.arange(n, **kw) - utils.ctrsub(n)) / n for n in sp.shape[1:]]
        grids = torch.meshgrid(*axes, indexing='ij')
        sel = sp.mask[0]
        for i in range(3):
            self.loc_[..., i] = self.fov[:, None, i] * grids[i][sel][None] + self.ofst[:, None, i]

    def applypulse(
        self, pulse: Pulse, *,
        doEmbed: bool = False, doRelax: bool = True, doUpdate: bool = False,
        b1Map: Optional[Tensor] = None, b1Map_: Optional[Tensor] = None
    ) -> Tensor:
        r"""Apply a pulse to the spincube object (uses its own ``loc_`` and ``Δf_``).

        Usage:
            ``M = spincube.applypulse(pulse, *, doEmbed=True, doRelax, doUpdate, b1Map)``
            ``M_ = spincube.applypulse(pulse, *, doEmbed=False, doRelax, doUpdate, b1Map_)``
        """
        assert ((b1Map_ is None) or (b1Map is None))
        b1Map_ = b1Map_ if b1Map is None else self.extract(b1Map)
        return self.spinarray.applypulse(pulse, doEmbed=doEmbed, doRelax=doRelax, doUpdate=doUpdate,
                                         Δf_=self.Δf_, loc_=self.loc_, b1Map_=b1Map_)

    def freeprec(self, dur: Tensor, *, doEmbed: bool = False, doRelax: bool = True,
                 doUpdate: bool = False) -> Tensor:
        r"""``M(_) = spincube.freeprec(dur, *, doEmbed, doRelax, doUpdate)`` with the cube's own ``Δf_``."""
        return self.spinarray.freeprec(dur, Δf_=self.Δf_, doEmbed=doEmbed, doRelax=doRelax, doUpdate=doUpdate)

    def asdict(self, *, toNumpy: bool = True, doEmbed: bool = True) -> dict:
        r"""``d = spincube.asdict(*, toNumpy, doEmbed)``: ``loc, Δf, fov, ofst`` + the SpinArray entries."""
        conv = (lambda x: x.detach().cpu().numpy()) if toNumpy else (lambda x: x.detach())
        d = {k: conv(getattr(self, k)) for k in ('loc', 'Δf')}
        d.update({k: getattr(self, k) for k in ('fov', 'ofst')})
        d.update(self.spinarray.asdict(toNumpy=toNumpy, doEmbed=doEmbed))
        return d

    def pulse2beff(self, pulse: Pulse, *, doEmbed: bool = False, b1Map: Optional[Tensor] = None,
                   b1Map_: Optional[Tensor] = None) -> Tensor:
        r"""``beff(_) = spincube.pulse2beff(pulse, *, doEmbed, b1Map ⊻ b1Map_)`` `(N,*Nd ⊻ nM,nT,xyz)`.
        (Upstream passes ``loc_`` positionally to a keyword-only parameter and raises, mobjs.py:942.)"""
        return self.spinarray.pulse2beff(pulse, loc_=self.loc_, doEmbed=doEmbed, Δf_=self.Δf_,
                                         b1Map=b1Map, b1Map_=b1Map_)

    def to(self, *, device: torch.device = torch.device('cpu'),
           dtype: torch.dtype = torch.float32) -> 'SpinCube':
        r"""The same cube on ``device`` with ``dtype`` (``self`` when nothing changes)."""
        if self.device == device and self.dtype == dtype:
            return self
        return SpinCube(self.shape, self.fov, mask=self.mask, ofst=self.ofst, Δf_=self.Δf_, T1_=self.T1_,
                        T2_=self.T2_, γ_=self.γ_, M_=self.M_, device=device, dtype=dtype)


class SpinBolus(SpinArray):
    """Placeholder, as upstream (mobjs.py:968-973)."""

    def __init__(self):
        pass


class Examples(object):
    r"""Quick exemplary instances to play around with (mobjs.py:976-1038)."""

    @staticmethod
    def pulse() -> Pulse:
        r"""A 512-sample pulse: circular 10 G rf, unit x/y gradients, arctan z-gradient."""
        kw = {'dtype': torch.float32, 'device': torch.device('cpu')}
        N, nT = 1, 512
        t = torch.arange(0, nT, **kw).reshape((N, 1, nT))
        rf = 10 * torch.cat([torch.cos(t / nT * 2 * π), torch.sin(t / nT * 2 * π)], 1)
        one = torch.ones((N, 1, nT), **kw)
        gr = torch.cat([one, one, 10 * torch.atan(t - round(nT / 2)) / π], 1)
        return Pulse(rf=rf, gr=gr, dt=dt0, **kw)

    @staticmethod
    def _mask333():
        mask = torch.zeros((1, 3, 3, 3), dtype=torch.bool)
        mask[0, :, 1, :], mask[0, 1, :, :] = True, True
        return mask

    @staticmethod
    def spinarray() -> SpinArray:
        r"""A masked 3×3×3 SpinArray (15 spins), T1 = 1 s, T2 = 40 ms."""
        kw = {'dtype': torch.float32, 'device': torch.device('cpu')}
        return SpinArray((1, 3, 3, 3), mask=Examples._mask333(), T1_=tensor([[1.]], **kw),
                         T2_=tensor([[4e-2]], **kw), γ_=γH, **kw)

    @staticmethod
    def spincube() -> SpinCube:
        r"""The same spins as a SpinCube with fov 3 cm, offset z = 1 cm and Δf cancelling unit x/y gradients."""
        kw = {'dtype': torch.float32, 'device': torch.device('cpu')}
        cube = SpinCube((1, 3, 3, 3), tensor([[3., 3., 3.]], **kw), mask=Examples._mask333(),
                        ofst=tensor([[0., 0., 1.]], **kw), T1_=tensor([[1.]], **kw), T2_=tensor([[4e-2]], **kw),
                        γ_=γH, **kw)
        cube.Δf = torch.sum(-cube.loc[0:1, :, :, :, 0:2], dim=-1) * γH
        return cube
