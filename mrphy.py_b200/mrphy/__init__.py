r"""mrphy -- B200-native drop-in for the differentiable Bloch-simulation path of MRphy.py.

Same public surface as the reference package (``/root/reference/mrphy/__init__.py``): the constants
below are float64 0-dim tensors and are the defaults of every keyword argument
(``__init__.py:58-65`` upstream), and the submodules keep their names:

    utils, beffective, sims, slowsims, mobjs

Naming conventions are the reference's: a trailing ``_`` marks a *compact* array of shape
``(N, nM, ...)`` instead of ``(N, *Nd, ...)``; ``N`` batch, ``nM`` spins, ``nT`` time points.

What differs: the simulation itself (``sims.blochsim`` / ``BlochSim``, ``SpinArray.applypulse``,
``beffective.rfgr2beff`` / ``beff2uφ`` / ``beff2ab``) runs as hand-written CUDA for sm_100a behind a
C ABI (``include/mrphy_b200.h``, ``libmrphy_b200.so``).  There is no CPU implementation of that
path: calling it with CPU tensors, or without the built library, raises ``RuntimeError``.
"""
import ctypes
from math import inf, pi as π  # noqa: F401

import torch
from torch import tensor


def _f64(value) -> torch.Tensor:
    """Package constants are float64 0-dim tensors: they are the default of every keyword argument of the API."""
    return tensor(value, dtype=torch.float64)


γH = _f64(4257.6)        # Hz/Gauss   gyromagnetic ratio of the water proton
T1G, T2G = _f64(1.47), _f64(0.07)                     # s          grey matter
dt0 = _f64(4e-6)         # s          default dwell time
gmax0, smax0, rfmax0 = _f64(5), _f64(12e3), _f64(0.25)   # Gauss/cm, Gauss/cm/s, Gauss: default hardware limits

_slice = slice(None)


def cuda_is_available() -> bool:
    r"""Whether a CUDA driver library can be dlopen-ed (the probe the reference uses, ``__init__.py:70-83``)."""
    def loads(name):
        try:
            ctypes.CDLL(name)
        except OSError:
            return False
        return True
    return any(loads(n) for n in ('libcuda.so', 'libcuda.so.1', 'libcuda.dylib', 'cuda.dll'))


__CUDA_IS_AVAILABLE__ = cuda_is_available()

try:   # cupy is only used by utils.rf_c2r / rf_r2c on cupy arrays
    import cupy  # noqa: F401
except ImportError:
    __CUPY_IS_AVAILABLE__ = False
else:
    __CUPY_IS_AVAILABLE__ = True

from mrphy import (utils, beffective, sims, slowsims, mobjs)  # noqa: E402
from mrphy.version import __version__  # noqa: E402,F401

__all__ = ['γH', 'utils', 'beffective', 'sims', 'slowsims', 'mobjs']
