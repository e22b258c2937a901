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

from math import pi as π, inf  # noqa: F401
import torch
from torch import tensor

γH = tensor(4257.6, dtype=torch.double)    # Hz/Gauss, water proton gyromagnetic ratio
T1G = tensor(1.47, dtype=torch.double)     # s, grey-matter T1
T2G = tensor(0.07, dtype=torch.double)     # s, grey-matter T2

dt0 = tensor(4e-6, dtype=torch.double)     # s, default dwell time
gmax0 = tensor(5, dtype=torch.double)      # Gauss/cm
smax0 = tensor(12e3, dtype=torch.double)   # Gauss/cm/s
rfmax0 = tensor(0.25, dtype=torch.double)  # Gauss

_slice = slice(None)


def cuda_is_available() -> bool:
    r"""``True`` when a CUDA driver library can be loaded (same probe as the reference)."""
    for name in ('libcuda.so', 'libcuda.so.1', 'libcuda.dylib', 'cuda.dll'):
        try:
            ctypes.CDLL(name)
            return True
        except OSError:
            pass
    return False


__CUDA_IS_AVAILABLE__ = cuda_is_available()

try:
    import cupy  # noqa: F401
    __CUPY_IS_AVAILABLE__ = True
except ImportError:
    __CUPY_IS_AVAILABLE__ = False

from mrphy import (utils, beffective, sims, slowsims, mobjs)  # noqa: E402
from mrphy.version import __version__  # noqa: E402,F401

__all__ = ['γH', 'utils', 'beffective', 'sims', 'slowsims', 'mobjs']
