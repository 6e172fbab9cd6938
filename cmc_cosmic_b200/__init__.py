"""cmc_cosmic_b200 -- B200-native sort / potential / new-orbit path of ClusterMonteCarlo/CMC-COSMIC.

The product is the C-ABI shared library ``libcmcb200.so`` (CUDA kernels for sm_100a, declared in
``include/cmcb200.h``).  This package is the host-side mirror of the reference's interface for that path
(same function names, argument meaning and error behaviour) used by the tests and by ``bench.py``.
There is no CPU fallback: importing :mod:`cmc_cosmic_b200.lib` without the built library raises.
"""
from .layout import STAR_DTYPE_SIZE, BINARY_DTYPE_SIZE, SF_INFINITY, ZERO  # noqa: F401

__all__ = ["lib", "cluster", "models", "layout", "partition"]
