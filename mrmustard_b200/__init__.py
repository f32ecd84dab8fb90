"""mrmustard_b200 — B200-native (sm_100a) Gaussian-to-Fock hot path for MrMustard.

Only the path of BASELINE.json's north_star lives here: the renormalized multidimensional-Hermite
recurrence (Bargmann triple -> Fock lattice) and its VJP, as hand-written CUDA kernels behind a C ABI
(include/mmhermite.h), plus the host-side mirror of the reference's operator interface:

  mrmustard_b200.strategies   — vanilla_numba, stable_numba, vanilla_batch_numba, vanilla_vjp_numba, ...
  mrmustard_b200.backend      — hermite_renormalized* with the BackendManager's batching semantics
  mrmustard_b200.dropin       — install() into a live `mrmustard.math`

Importing the package loads the CUDA library; there is no CPU fallback.
"""
from . import _lib, strategies  # noqa: F401
from .backend import (  # noqa: F401
    hermite_renormalized,
    hermite_renormalized_batched,
    hermite_renormalized_1leftoverMode,
    hermite_renormalized_binomial,
    hermite_renormalized_contracted,
    hermite_renormalized_diagonal,
)

__version__ = "0.1.0"
