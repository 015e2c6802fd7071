"""B200-native incompressible Navier-Stokes time step: drop-in for the solver entry points of
mhw32/neural-navier-stokes (``src/chorin_fd``, ``src/direct_fd``, ``src/chorin_spectral``,
``src/boundary.py``, ``src/constants.py``).

Host code is Python; all numerics run in hand-written sm_100a CUDA kernels behind the C ABI of
``include/nns_b200.h`` (``libnns_b200.so``, loaded with ctypes).  There is no CPU fallback:
creating a solver without the built library or without a CUDA device raises.

Import as ``nns_b200`` (alias module at the repo root).
"""
from . import boundary, constants  # noqa: F401
from .boundary import (BaseBoundaryCondition, DirichletBoundaryCondition,  # noqa: F401
                       NeumannBoundaryCondition)

__all__ = ["boundary", "constants", "BaseBoundaryCondition", "DirichletBoundaryCondition",
           "NeumannBoundaryCondition"]
